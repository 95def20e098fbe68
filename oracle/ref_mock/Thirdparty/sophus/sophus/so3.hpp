// ORACLE — TEST INFRASTRUCTURE ONLY.  Mock of the Sophus header the reference's Common.h includes; see mock_eigen.hpp.
#pragma once
#include "../../../mock_eigen.hpp"
