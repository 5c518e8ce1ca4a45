// forwarding header: same include name as the reference (hpipm-cpp/include/hpipm-cpp/ocp_qp_ipm_solver.hpp)
#pragma once
#include "hpipm-cpp.hpp"
