// merger.cuh -- reduce-side k-way merge on device (placeholder until the merge kernels land).
#pragma once
#include "sorter.cuh"
