// ORACLE — TEST INFRASTRUCTURE ONLY.  Mock of the OpenCV header the reference includes; see mock_cv.hpp.
#pragma once
#include "../../mock_cv.hpp"
