// TEST INFRASTRUCTURE: see ../SpaceInformation.h
#pragma once
#include "ompl/base/SpaceInformation.h"
