// Microbenchmark (GPU box): throughput of RANDOM accesses to a large HBM-resident buffer, by access
// granularity -- what the per-query search state of k_astar_lane does (node records, node table,
// heap tail: 32 B sectors scattered over ~15 GB).  Prints GB/s of useful bytes per pattern.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o randmem randmem.cu && ./randmem [GiB]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

// each thread: ITERS rounds of MLP independent accesses of GRAN bytes (GRAN/16 uint4 loads each)
// regionGran > 0: thread t only touches granules [t * regionGran, (t + 1) * regionGran) -- the layout of
// k_astar_lane's per-lane state (a warp's 32 accesses fall into a few 2 MB pages); 0: anywhere in the buffer
template <int GRAN, int MLP, bool WRITE, int WBYTES>
__global__ void k_rand(uint4* buf, uint64_t nGran, uint32_t regionGran, int iters, uint64_t seed, uint32_t* sink) {
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    uint64_t idx[MLP];
#pragma unroll
    for (int m = 0; m < MLP; ++m) {
      const uint64_t r = mix(seed + tid * 1315423911ULL + (uint64_t)it * MLP + m);
      idx[m] = regionGran ? tid * regionGran + (uint32_t)(r >> 32) % regionGran : r % nGran;
    }
    if (WRITE) {
#pragma unroll
      for (int m = 0; m < MLP; ++m) {
        char* p = reinterpret_cast<char*>(buf) + idx[m] * GRAN;
        if (WBYTES == 2) *reinterpret_cast<uint16_t*>(p + 2 * (tid & 7)) = (uint16_t)it;
        else if (WBYTES == 4) *reinterpret_cast<uint32_t*>(p + 4 * (tid & 3)) = it;
        else if (WBYTES == 16) *reinterpret_cast<uint4*>(p) = make_uint4(it, it, it, it);
        else {
#pragma unroll
          for (int k = 0; k < GRAN / 16; ++k) reinterpret_cast<uint4*>(p)[k] = make_uint4(it, k, it, k);
        }
      }
    } else {
      uint4 v[MLP][GRAN / 16];
#pragma unroll
      for (int m = 0; m < MLP; ++m)
#pragma unroll
        for (int k = 0; k < GRAN / 16; ++k) v[m][k] = __ldcg(&buf[idx[m] * (GRAN / 16) + k]);
#pragma unroll
      for (int m = 0; m < MLP; ++m)
#pragma unroll
        for (int k = 0; k < GRAN / 16; ++k) acc += v[m][k].x ^ v[m][k].w;
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

// dependent chain: each thread chases pointers (MLP = 1 per thread); latency under load
__global__ void k_chase(const uint4* buf, uint64_t nGran, int iters, uint64_t seed, uint32_t* sink) {
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t x = mix(seed + tid);
  for (int it = 0; it < iters; ++it) {
    const uint4 v = __ldcg(&buf[(x % nGran) * 2]);
    x = mix(x + v.x + it);
  }
  if (x == 0x12345678u) *sink = (uint32_t)x;
}

template <class F>
float timeit(F&& f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

template <int GRAN, int MLP, bool WRITE, int WBYTES>
void run(const char* name, uint4* buf, uint64_t bytes, int warpsPerSm, uint32_t* sink, bool regions = false) {
  const int blocks = 148 * warpsPerSm, threads = 32, iters = 400;
  const uint64_t nGran = bytes / GRAN;
  const uint32_t regionGran = regions ? (uint32_t)(nGran / ((uint64_t)blocks * threads)) : 0u;
  float ms = timeit([&] { k_rand<GRAN, MLP, WRITE, WBYTES><<<blocks, threads>>>(buf, nGran, regionGran, iters, 12345, sink); });
  if (regions) printf("[per-thread regions of %u KB] ", regionGran * GRAN / 1024);
  CK(cudaGetLastError());
  const double acc = (double)blocks * threads * iters * MLP;
  const int useful = WRITE && WBYTES ? WBYTES : GRAN;
  printf("%-34s foot %5.1f GiB warps/SM %2d MLP %d: %7.2f G acc/s  useful %7.1f GB/s  sectors(32B) %7.1f GB/s\n", name, bytes / 1073741824.0,
         warpsPerSm, MLP, acc / ms * 1e-6, acc * useful / ms * 1e-6, acc * (GRAN < 32 ? 32 : GRAN) / ms * 1e-6);
}

int main(int argc, char** argv) {
  const double gib = argc > 1 ? atof(argv[1]) : 16.0;
  const uint64_t bytes = (uint64_t)(gib * 1073741824.0) & ~0xfffULL;
  uint4* buf;
  uint32_t* sink;
  CK(cudaMalloc(&buf, bytes));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(buf, 1, bytes));
  // the access pattern of k_astar_lane: every thread inside its own region
  for (int w : {8, 16, 32}) {
    run<32, 1, false, 0>("read 32 B", buf, bytes, w, sink, true);
    run<32, 4, false, 0>("read 32 B", buf, bytes, w, sink, true);
    run<64, 4, false, 0>("read 64 B", buf, bytes, w, sink, true);
    run<128, 4, false, 0>("read 128 B", buf, bytes, w, sink, true);
    run<32, 4, true, 0>("write 32 B (full sector)", buf, bytes, w, sink, true);
    run<32, 4, true, 16>("write 16 B (half sector)", buf, bytes, w, sink, true);
    run<32, 4, true, 2>("write 2 B (partial)", buf, bytes, w, sink, true);
  }
  for (uint64_t foot : {bytes, (uint64_t)64 << 20}) {
    if (foot > bytes) continue;
    for (int w : {16, 32, 64}) {
      run<32, 1, false, 0>("read 32 B", buf, foot, w, sink);
      run<32, 4, false, 0>("read 32 B", buf, foot, w, sink);
    }
    run<64, 4, false, 0>("read 64 B", buf, foot, 32, sink);
    run<128, 4, false, 0>("read 128 B", buf, foot, 32, sink);
    run<256, 2, false, 0>("read 256 B", buf, foot, 32, sink);
    run<32, 4, true, 0>("write 32 B (full sector)", buf, foot, 32, sink);
    run<32, 4, true, 16>("write 16 B (half sector)", buf, foot, 32, sink);
    run<32, 4, true, 4>("write 4 B (partial)", buf, foot, 32, sink);
    run<32, 4, true, 2>("write 2 B (partial)", buf, foot, 32, sink);
    run<128, 4, true, 0>("write 128 B", buf, foot, 32, sink);
    for (int w : {16, 32, 64}) {
      const int blocks = 148 * w, iters = 300;
      float ms = timeit([&] { k_chase<<<blocks, 32>>>(buf, foot / 32, iters, 99, sink); });
      printf("chase 32 B dependent  foot %5.1f GiB warps/SM %2d: %.2f G acc/s, %.0f ns per hop per thread\n", foot / 1073741824.0, w,
             (double)blocks * 32 * iters / ms * 1e-6, ms * 1e6 / iters);
    }
  }
  return 0;
}
