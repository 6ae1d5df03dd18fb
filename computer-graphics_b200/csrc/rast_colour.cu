// Colour modes 1 / 2 of the rasteriser (rasteriser/Source/skeleton.cpp:647-662): the keys of the accepted
// fragments (triangle << 24 | y << 12 | x) sorted into the serial order of the reference's loops, so that
// the n-th accepted fragment gets the n-th triple of rand() values.  The sort is CUB's device radix sort --
// a library call, used by this toy mode only (never on a benchmarked path).
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

int rast_colour_sort(b200_ctx *ctx, const unsigned long long *in, unsigned long long *out, unsigned long long n, int end_bit) {
  size_t tmp_bytes = 0;
  CU_CHECK(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, in, out, (int)n, 0, end_bit, ctx->stream));
  if (int rc = ensure(ctx, ctx->rast_colour_tmp, tmp_bytes ? tmp_bytes : 1)) return rc;
  CU_CHECK(ctx, cub::DeviceRadixSort::SortKeys(ctx->rast_colour_tmp.p, tmp_bytes, in, out, (int)n, 0, end_bit, ctx->stream));
  ctx->stats.kernel_launches++;
  return B200_OK;
}
