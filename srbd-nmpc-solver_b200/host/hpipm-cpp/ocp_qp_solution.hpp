// forwarding header: same include name as the reference (hpipm-cpp/include/hpipm-cpp/ocp_qp_solution.hpp)
#pragma once
#include "hpipm-cpp.hpp"
