// nvtx.cuh -- NVTX ranges around solver calls and their batches (tracing row: CUDA events + NVTX + ncu).
// NVTX v3 is header-only: without a profiler attached a range costs one predictable branch.
#pragma once
#include <nvtx3/nvToolsExt.h>

struct VclRange
{
  explicit VclRange(const char *name) { nvtxRangePushA(name); }
  ~VclRange() { nvtxRangePop(); }
  VclRange(const VclRange&) = delete;
  VclRange &operator=(const VclRange&) = delete;
};
#define VCL_RANGE_CAT2(a, b) a##b
#define VCL_RANGE_CAT(a, b) VCL_RANGE_CAT2(a, b)
#define VCL_RANGE(name) VclRange VCL_RANGE_CAT(vcl_range_, __LINE__)(name)
