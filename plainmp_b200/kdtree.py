"""Host mirror of ``plainmp.kdtree`` (reference: cpp/plainmp/bindings/kdtree.cpp:19-23,
cpp/plainmp/collision/kdtree.cpp:20-96).

The reference builds a median-split kd-tree and answers one exact nearest-neighbour query at a
time by recursive descent.  The layout of that tree is not observable through ``query`` /
``sqdist``; here the points are binned into the device's uniform grid (csrc/pmp_cloud.cuh) and
queries are answered exactly on the GPU, in batches.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from . import tables
from .psdf import CloudSDF


class KDTree:
    def __init__(self, points: np.ndarray):
        self._points = np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(-1, 3))
        self._scene = tables.SceneTable([CloudSDF(self._points, 0.0)])

    def sqdist_batch(self, targets: np.ndarray) -> np.ndarray:
        return self._scene.cloud_nearest(np.asarray(targets, dtype=np.float64).reshape(-1, 3))[0]

    def query_batch(self, targets: np.ndarray) -> np.ndarray:
        idx = self._scene.cloud_nearest(np.asarray(targets, dtype=np.float64).reshape(-1, 3))[1]
        return self._points[idx]

    def query(self, target: Sequence[float]) -> np.ndarray:
        return self.query_batch(np.asarray(target).reshape(1, 3))[0]

    def sqdist(self, target: Sequence[float]) -> float:
        return float(self.sqdist_batch(np.asarray(target).reshape(1, 3))[0])
