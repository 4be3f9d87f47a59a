// TEST INFRASTRUCTURE: see ../cuda_emul.h
#pragma once
#include "cuda_emul.h"
