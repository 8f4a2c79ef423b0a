#include "common.cuh"
