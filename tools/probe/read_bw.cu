// read_bw.cu — read-only HBM bandwidth probe for the access patterns the scan kernels can use.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o read_bw read_bw.cu && ./read_bw [GiB]
// Variants: grid-stride LDG.128 (different unroll / occupancy), warp-private regions with a 4-slot register
// ring (the scan kernels' pattern), and a per-warp cp.async.bulk (TMA engine, no tensor map) shared-memory ring.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int U>
__global__ void gridstride(const uint4* __restrict__ p, uint64_t nvec, unsigned long long* out)
{
  uint32_t acc = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < nvec; i += U * stride)
  {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(p + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  for (; i < nvec; i += stride) { uint4 v = __ldg(p + i); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0x12345678u) atomicAdd(out, 1ull);
}

struct Sched {
  unsigned long long* ticket; uint64_t tw, w, rounds, st, i;
  __device__ void init(unsigned long long* t, uint64_t nreg) {
    ticket = t; tw = (uint64_t)gridDim.x * (blockDim.x >> 5); w = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    rounds = (nreg - nreg / 8) / tw; st = rounds * tw; i = 0; }
  __device__ uint64_t next(uint32_t lane) {
    if (i < rounds) return (i++) * tw + w;
    unsigned long long tk = 0; if (lane == 0) tk = atomicAdd(ticket, 1ull);
    return st + __shfl_sync(0xffffffffu, tk, 0); }
};

// fake per-chunk work: about WORK dependent-free ALU ops on the 16 bytes
template <int WORK>
__device__ __forceinline__ uint32_t chew(uint4 v, uint32_t acc)
{
  uint32_t a = v.x, b = v.y, c = v.z, d = v.w;
#pragma unroll
  for (int k = 0; k < WORK / 8; ++k)
  {
    a = (a ^ 0x53535353u) - 0x01010101u; b = (b ^ 0x6c6c6c6cu) - 0x01010101u;
    c = (c | a) ^ (b & 0x80808080u);     d = (d | b) ^ (a & 0x80808080u);
  }
  return acc ^ a ^ b ^ c ^ d;
}

// warp-private 16 KiB regions, 4-slot ring of LDG.128 (2 KiB in flight per warp)
template <int REGION, int WORK>
__global__ void warpring(const uint8_t* __restrict__ buf, uint64_t n, unsigned long long* ticket, unsigned long long* out)
{
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t nreg = n / REGION;
  uint32_t acc = 0;
  Sched sc; sc.init(ticket, nreg);
  uint64_t r = sc.next(lane);
  uint4 v[4];
  if (r < nreg)
    for (int j = 0; j < 4; ++j) v[j] = __ldg((const uint4*)(buf + r * REGION + j * 512 + lane * 16));
  while (r < nreg)
  {
    const uint64_t rn = sc.next(lane);
    const uint8_t* p = buf + r * REGION + lane * 16;
    for (int b = 0; b < REGION / 2048; ++b)
    {
      const bool last = b + 1 == REGION / 2048;
      const bool has = !last || rn < nreg;
      const uint8_t* pn = last ? buf + rn * REGION + lane * 16 : p + 2048;
#pragma unroll
      for (int j = 0; j < 4; ++j)
      {
        acc = chew<WORK>(v[j], acc);
        if (has) v[j] = __ldg((const uint4*)(pn + j * 512));
      }
      p += 2048;
    }
    r = rn;
  }
  if (acc == 0x12345678u) atomicAdd(out, 1ull);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// per-warp bulk-copy ring: STAGES x BYTES per warp; lane 0 issues cp.async.bulk, all lanes read with LDS.128
template <int STAGES, int BYTES, int WORK>
__global__ void bulkring(const uint8_t* __restrict__ buf, uint64_t n, unsigned long long* ticket, unsigned long long* out)
{
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint8_t* ring = smem + (size_t)wid * STAGES * BYTES;
  uint64_t* bars = (uint64_t*)(smem + (size_t)nw * STAGES * BYTES) + wid * STAGES;
  constexpr int REGION = 16384;
  constexpr int PER = REGION / BYTES;
  if (lane == 0)
    for (int s = 0; s < STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + s)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const uint64_t nreg = n / REGION;
  uint32_t acc = 0;
  // flat sequence of chunks: (region, piece); keep STAGES-1 in flight
  Sched sc; sc.init(ticket, nreg);
  uint64_t r_issue = sc.next(lane);
  uint32_t piece_issue = 0;
  uint64_t r_cons = r_issue;
  uint32_t piece_cons = 0;
  uint32_t issued = 0, consumed = 0;
  auto issue = [&]() {
    if (r_issue >= nreg) return;
    const uint32_t s = issued % STAGES;
    if (lane == 0)
    {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bars + s)), "r"(BYTES) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(ring + s * BYTES)), "l"(buf + r_issue * REGION + (uint64_t)piece_issue * BYTES), "r"(BYTES),
                   "r"(smem_u32(bars + s)) : "memory");
    }
    ++issued;
    if (++piece_issue == PER)
    {
      piece_issue = 0;
      r_issue = sc.next(lane);
    }
  };
  for (int s = 0; s < STAGES - 1; ++s) issue();
  while (consumed < issued)
  {
    issue();
    const uint32_t s = consumed % STAGES;
    const uint32_t parity = (consumed / STAGES) & 1;
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(smem_u32(bars + s)), "r"(parity) : "memory");
    const uint8_t* src = ring + s * BYTES;
#pragma unroll
    for (int k = 0; k < BYTES / 512; ++k)
    {
      uint4 v = *(const uint4*)(src + k * 512 + lane * 16);
      acc = chew<WORK>(v, acc);
    }
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ++consumed;
    (void)r_cons; (void)piece_cons;
  }
  if (acc == 0x12345678u) atomicAdd(out, 1ull);
}

int main(int argc, char** argv)
{
  const double gib = argc > 1 ? atof(argv[1]) : 4.0;
  const uint64_t n = (uint64_t)(gib * (1ull << 30)) / (1 << 20) * (1 << 20);
  uint8_t* buf; unsigned long long* out;
  CK(cudaMalloc(&buf, n + 4096));
  CK(cudaMemset(buf, 0x61, n + 4096));
  CK(cudaMalloc(&out, 64));
  CK(cudaMemset(out, 0, 64));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  auto timeit = [&](const char* name, auto launch) {
    float best = 1e9;
    for (int it = 0; it < 6; ++it)
    {
      CK(cudaMemset(out + 1, 0, 8));
      CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    printf("%-34s %8.3f ms  %8.1f GB/s\n", name, best, n / best / 1e6);
  };
  const uint64_t nvec = n / 16;
  timeit("gridstride U4 t256 x8/SM", [&] { gridstride<4><<<sms * 8, 256>>>((const uint4*)buf, nvec, out); });
  timeit("gridstride U4 t256 x4/SM", [&] { gridstride<4><<<sms * 4, 256>>>((const uint4*)buf, nvec, out); });
  timeit("gridstride U8 t256 x4/SM", [&] { gridstride<8><<<sms * 4, 256>>>((const uint4*)buf, nvec, out); });
  timeit("gridstride U8 t512 x4/SM", [&] { gridstride<8><<<sms * 4, 512>>>((const uint4*)buf, nvec, out); });
  timeit("gridstride U2 t1024 x2/SM", [&] { gridstride<2><<<sms * 2, 1024>>>((const uint4*)buf, nvec, out); });
  timeit("gridstride U1 t1024 x2/SM", [&] { gridstride<1><<<sms * 2, 1024>>>((const uint4*)buf, nvec, out); });
  timeit("gridstride U4 many CTAs", [&] { gridstride<4><<<(unsigned)(nvec / 256 / 4), 256>>>((const uint4*)buf, nvec, out); });

#define WR(REG, WORK, PER) timeit("warpring " #REG " work" #WORK " x" #PER "/SM", [&] { warpring<REG, WORK><<<sms * PER, 256>>>(buf, n, out + 1, out); })
  WR(16384, 0, 4); WR(16384, 0, 8); WR(16384, 40, 4); WR(16384, 40, 8); WR(16384, 80, 4); WR(16384, 80, 8); WR(65536, 40, 4);
#define BR(ST, BY, WORK, PER) do { auto k = bulkring<ST, BY, WORK>; const int smem = 8 * ST * BY + 8 * ST * 8; \
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    timeit("bulkring " #ST "x" #BY " work" #WORK " x" #PER "/SM", [&] { k<<<sms * PER, 256, smem>>>(buf, n, out + 1, out); }); } while (0)
  BR(4, 2048, 0, 3); BR(4, 2048, 40, 3); BR(4, 2048, 80, 3); BR(3, 2048, 40, 4); BR(6, 2048, 40, 2); BR(8, 1024, 40, 3); BR(3, 4096, 40, 2);
  BR(2, 2048, 40, 6); BR(4, 1024, 40, 6);
  // plain device-to-device copy for reference (read + write bytes)
  {
    uint8_t* dst; CK(cudaMalloc(&dst, n / 2));
    float best = 1e9;
    for (int it = 0; it < 5; ++it)
    {
      CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(dst, buf, n / 2, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    printf("%-34s %8.3f ms  %8.1f GB/s (read+write)\n", "cudaMemcpy D2D", best, n / best / 1e6);
  }
  return 0;
}
