// TEST INFRASTRUCTURE ONLY: see ../core/core.hpp
#pragma once
#include "../core/core.hpp"
