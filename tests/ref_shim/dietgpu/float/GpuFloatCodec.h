// shim: the reference's tests include this name; the API comes from the product's drop-in layer
#pragma once
#include "dietgpu_b200_compat.hpp"
