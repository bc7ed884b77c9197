// Maps of the fused one-ring halo exchange (include/zpcb200.h: zpc_halo_view) — built on the device from the gathered block codes of
// every rank, nothing read back to the host.  The exchange itself lives in the binned P2G's write-back (send) and in the grid update
// (receive): csrc/mpm_binned.cu halo_send_tile, csrc/mpm.cu grid_update_bc_kernel.
#include <climits>

#include "common.cuh"

namespace {

constexpr long long CODE_MAX = LLONG_MAX;
constexpr long long BIAS = 1 << 20;  // block coordinates in [-2^20, 2^20)

__global__ void halo_codes_kernel(const int *__restrict__ keys, const int *__restrict__ cnt, int cap, long long *__restrict__ codes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  long long c = CODE_MAX;
  if (i < *cnt) c = (((long long)keys[3 * i] + BIAS) << 42) | (((long long)keys[3 * i + 1] + BIAS) << 21) | ((long long)keys[3 * i + 2] + BIAS);
  codes[i] = c;
}

// hit[r][b] = 1 if my block b is also active on rank r (binary search in r's ascending codes)
__global__ void halo_hit_kernel(const long long *__restrict__ all, int world, int rank, int cap, int *__restrict__ hit) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (b >= cap) return;
  const long long c = all[(size_t)rank * cap + b];
  int h = 0;
  if (c != CODE_MAX && r != rank) {
    const long long *row = all + (size_t)r * cap;
    int lo = 0, hi = cap;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (row[mid] < c) lo = mid + 1;
      else hi = mid;
    }
    h = lo < cap && row[lo] == c;
  }
  hit[(size_t)r * cap + b] = h;
}

__global__ void halo_compose_kernel(const int *__restrict__ hit, const int *__restrict__ pos, int world, int cap, int seg, int *__restrict__ peer_out,
                                    int *__restrict__ pos_out, int *status) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= cap) return;
  int k = 0;
  for (int r = 0; r < world; ++r)
    if (hit[(size_t)r * cap + b]) {
      const int p = pos[(size_t)r * cap + b];
      if (p >= seg) { if (status) atomicOr(status, ZPC_HALO_SEGMENT_FULL); continue; }
      if (k >= ZPCB200_HALO_K) { if (status) atomicOr(status, ZPC_HALO_TOO_MANY_PEERS); continue; }
      peer_out[(size_t)b * ZPCB200_HALO_K + k] = r;
      pos_out[(size_t)b * ZPCB200_HALO_K + k] = p;
      ++k;
    }
  for (; k < ZPCB200_HALO_K; ++k) { peer_out[(size_t)b * ZPCB200_HALO_K + k] = -1; pos_out[(size_t)b * ZPCB200_HALO_K + k] = 0; }
}

}  // namespace

extern "C" {

int zpcb200_halo_codes(zpc_hashtable_view tb, int capacity, long long *codes, zpc_stream_t stream) {
  if (!tb.activeKeys || !tb.cnt || !codes || capacity <= 0) return ZPCB200_E_BADARG;
  halo_codes_kernel<<<(capacity + 255) / 256, 256, 0, (cudaStream_t)stream>>>(tb.activeKeys, tb.cnt, capacity, codes);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_halo_build(void *temp, size_t *temp_bytes, const long long *all_codes, int world, int rank, int capacity, int seg, int *peer_out,
                       int *pos_out, int *status, zpc_stream_t stream) {
  if (!temp_bytes || world < 1 || world > ZPCB200_HALO_MAX_PEERS || (unsigned)rank >= (unsigned)world || capacity <= 0 || seg <= 0)
    return ZPCB200_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  size_t scan_bytes = 0;
  zpc_port none = {nullptr, 0, 0, 0, 1};
  int rc = zpcb200_exclusive_scan_sum_i32(nullptr, &scan_bytes, none, none, (size_t)capacity, nullptr);
  if (rc) return rc;
  const size_t arr = zpc_align_up(sizeof(int) * (size_t)world * capacity, 256);
  const size_t need = 2 * arr + zpc_align_up(scan_bytes, 256);
  if (!temp) { *temp_bytes = need; return ZPCB200_OK; }
  if (*temp_bytes < need) return ZPCB200_E_TEMP_TOO_SMALL;
  if (!all_codes || !peer_out || !pos_out) return ZPCB200_E_BADARG;
  int *hit = (int *)temp, *pos = (int *)((char *)temp + arr);
  void *scan_tmp = (char *)temp + 2 * arr;
  dim3 grid((capacity + 255) / 256, world);
  halo_hit_kernel<<<grid, 256, 0, s>>>(all_codes, world, rank, capacity, hit);
  ZPC_CHECK_LAUNCH();
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    zpc_port pi = {hit + (size_t)r * capacity, 0, 0, 0, 1}, po = {pos + (size_t)r * capacity, 0, 0, 0, 1};
    size_t sb = scan_bytes;
    rc = zpcb200_exclusive_scan_sum_i32(scan_tmp, &sb, pi, po, (size_t)capacity, s);
    if (rc) return rc;
  }
  halo_compose_kernel<<<(capacity + 255) / 256, 256, 0, s>>>(hit, pos, world, capacity, seg, peer_out, pos_out, status);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}
}
