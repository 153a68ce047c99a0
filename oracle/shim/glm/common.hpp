// TEST INFRASTRUCTURE -- see glm.hpp in this directory (stand-in for glm 0.9.9.8 sub-header).
#pragma once
#include "glm.hpp"
