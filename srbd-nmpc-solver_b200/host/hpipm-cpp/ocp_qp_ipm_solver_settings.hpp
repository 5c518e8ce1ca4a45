// forwarding header: same include name as the reference (hpipm-cpp/include/hpipm-cpp/ocp_qp_ipm_solver_settings.hpp)
#pragma once
#include "hpipm-cpp.hpp"
