// forwarding header: same include name as the reference (hpipm-cpp/include/hpipm-cpp/ocp_qp_dim.hpp)
#pragma once
#include "hpipm-cpp.hpp"
