// ggr_kernels_coop_enc.cu - lock-step request-side pass A (one warp per item); see ggr_coop_enc.cuh.
#include "ggr_kernels.h"
#include "ggr_coop_enc.cuh"

#define CE_WARPS 4

// Every item the lock-step parser handles gets size / first / status written here; the others are
// appended to `pending` (order irrelevant) for the per-thread parser.
__global__ void __launch_bounds__(CE_WARPS * 32)
k_encode_coop_parse(const u8* __restrict__ blob, long long n, u32 n_msgs, const i32* __restrict__ msg_id,
                    const u8* __restrict__ in, const u64* __restrict__ in_off, u8* __restrict__ ir, u32* __restrict__ size,
                    u32* __restrict__ first, i32* __restrict__ status, u32* __restrict__ pending, u32* __restrict__ n_pending) {
  __shared__ CoopEnc S[CE_WARPS];
  __shared__ u32 lut[256];
  for (u32 b = threadIdx.x; b < 256; b += CE_WARPS * 32) lut[b] = ce_class(b);
  __syncthreads();
  const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * CE_WARPS + warp;
  if (item >= n) return;
  const u64 a = in_off[item], b = in_off[item + 1];
  const i32 m = msg_id[item];
  bool ok = false;
  EncResult res;
  res.size = 0;
  res.first = GGR_NIL;
  if (m >= 0 && (u32)m < n_msgs && b >= a && b - a <= (u64)CE_MAX_INPUT - 16u) {
    const Tables T = ggr_tables(blob);
    const u64 node_off = (a >> 1) + 8ull * (u64)item;
    const u32 cap = (u32)(((b >> 1) + 8ull * (u64)(item + 1)) - node_off);
    const u8* base = in + (a & ~15ull);
    const u32 s0 = (u32)(a & 15ull);
    ok = ce_parse_item(S[warp], lut, T, (u32)m, base, s0, s0 + (u32)(b - a), ir + node_off * 16, cap, &res);
  }
  if (lane == 0) {
    if (ok) {
      size[item] = res.size;
      first[item] = res.first;
      status[item] = GST_OK;
    } else {
      size[item] = 0;
      pending[atomicAdd(n_pending, 1u)] = (u32)item;
    }
  }
}

void ggr_launch_encode_coop_parse(cudaStream_t st, long long n, const uint8_t* blob, uint32_t n_msgs, const int32_t* msg_id,
                                  const uint8_t* in, const uint64_t* in_off, uint8_t* ir, uint32_t* size, uint32_t* first,
                                  int32_t* status, uint32_t* pending, uint32_t* n_pending) {
  unsigned nb = (unsigned)((n + CE_WARPS - 1) / CE_WARPS);
  k_encode_coop_parse<<<nb, CE_WARPS * 32, 0, st>>>(blob, n, n_msgs, msg_id, in, (const u64*)in_off, ir, size, first, status,
                                                    pending, n_pending);
}
