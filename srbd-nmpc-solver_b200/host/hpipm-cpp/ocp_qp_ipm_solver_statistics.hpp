// forwarding header: same include name as the reference (hpipm-cpp/include/hpipm-cpp/ocp_qp_ipm_solver_statistics.hpp)
#pragma once
#include "hpipm-cpp.hpp"
