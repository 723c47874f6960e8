// ggr_kernels_coop.cu - warp-cooperative reply-side kernels (one warp per item); see ggr_coop.cuh.
#include "ggr_kernels.h"
#include "ggr_coop.cuh"

#define COOP_WARPS 4

__global__ void __launch_bounds__(COOP_WARPS * 32)
k_decode_coop_size(const u8* __restrict__ blob, long long n, u32 n_msgs, const i32* __restrict__ msg_id,
                   const u8* __restrict__ in, const u64* __restrict__ in_off, u32 flags, u32* __restrict__ size,
                   u32* __restrict__ mode, i32* __restrict__ status) {
  __shared__ CoopShared S[COOP_WARPS];
  const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long item = (long long)blockIdx.x * COOP_WARPS + warp;
  if (item >= n) return;
  u64 a = in_off[item], b = in_off[item + 1];
  i32 m = msg_id[item];
  bool ok = false;
  u32 sz = 0;
  if (m >= 0 && (u32)m < n_msgs && b >= a && b - a <= 0x7FFFFFF0ull) {
    DecCtx cx;
    cx.T = ggr_tables(blob);
    cx.in = in + (a & ~15ull);
    cx.flags = flags;
    u32 s0 = (u32)(a & 15ull);
    int ws = GST_OK;
    ok = coop_decode_item(S[warp], cx, (u32)m, s0, s0 + (u32)(b - a), lane, 32, nullptr, 0, &sz, &ws);
  }
  if (lane == 0) {
    if (ok) {
      size[item] = sz;
      mode[item] = GGR_MODE_COOP;
      status[item] = GST_OK;
    } else {
      mode[item] = GGR_MODE_PENDING;
    }
  }
}

__global__ void __launch_bounds__(COOP_WARPS * 32)
k_decode_coop_write(const u8* __restrict__ blob, long long n, const i32* __restrict__ msg_id, const u8* __restrict__ in,
                    const u64* __restrict__ in_off, u32 flags, const u32* __restrict__ size, const u32* __restrict__ mode,
                    i32* __restrict__ status, u8* __restrict__ out, const u64* __restrict__ out_off) {
  __shared__ CoopShared S[COOP_WARPS];
  const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long item = (long long)blockIdx.x * COOP_WARPS + warp;
  if (item >= n) return;
  if (mode[item] != GGR_MODE_COOP || status[item] != GST_OK) return;
  u64 a = in_off[item], b = in_off[item + 1];
  u64 goff = out_off[item];
  DecCtx cx;
  cx.T = ggr_tables(blob);
  cx.in = in + (a & ~15ull);
  cx.flags = flags;
  u32 s0 = (u32)(a & 15ull);
  u32 sz = 0;
  int ws = GST_OK;
  bool ok = coop_decode_item(S[warp], cx, (u32)msg_id[item], s0, s0 + (u32)(b - a), lane, 32, out + (goff & ~7ull),
                             (u32)(goff & 7ull), &sz, &ws);
  bool bad = !ok || ws != GST_OK || sz != size[item];
  if (__any_sync(0xFFFFFFFFu, bad) && lane == 0) status[item] = GST_INTERNAL;
}

void ggr_launch_decode_coop_size(cudaStream_t st, long long n, const uint8_t* blob, uint32_t n_msgs, const int32_t* msg_id,
                                 const uint8_t* in, const uint64_t* in_off, uint32_t flags, uint32_t* size, uint32_t* mode,
                                 int32_t* status) {
  unsigned nb = (unsigned)((n + COOP_WARPS - 1) / COOP_WARPS);
  k_decode_coop_size<<<nb, COOP_WARPS * 32, 0, st>>>(blob, n, n_msgs, msg_id, in, (const u64*)in_off, flags, size, mode, status);
}
void ggr_launch_decode_coop_write(cudaStream_t st, long long n, const uint8_t* blob, const int32_t* msg_id, const uint8_t* in,
                                  const uint64_t* in_off, uint32_t flags, const uint32_t* size, const uint32_t* mode,
                                  int32_t* status, uint8_t* out, const uint64_t* out_off) {
  unsigned nb = (unsigned)((n + COOP_WARPS - 1) / COOP_WARPS);
  k_decode_coop_write<<<nb, COOP_WARPS * 32, 0, st>>>(blob, n, msg_id, in, (const u64*)in_off, flags, size, mode, status, out,
                                                     (const u64*)out_off);
}
