// the reference's tests include <sycl.hpp> (DPC++ spelling); same header
#pragma once
#include "sycl/sycl.hpp"
