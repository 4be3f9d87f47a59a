// TEST INFRASTRUCTURE: constraints/composite_constraint.hpp includes the pybind11 headers without using them;
// the reference's hot-path sources are compiled here without Python (oracle/Makefile.ref).
#pragma once
