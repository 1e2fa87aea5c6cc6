// Ceiling probe for the SpMM access pattern: random gathers of whole embedding rows (ROWB bytes) from a
// table that fits in L2, nothing else (no index decode beyond one coalesced int load, no FMAs kept).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/l2_gather_probe.cu -o /tmp/l2probe && /tmp/l2probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int LPR, int U>   // LPR lanes per row (each a float4), U rows in flight per lane group
__global__ void __launch_bounds__(256) probe(const float4* __restrict__ X, const int* __restrict__ idx, long n_gather,
                                             float* __restrict__ out) {
  const int lig = threadIdx.x % LPR;
  const long group = ((long)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const long n_groups = ((long)gridDim.x * blockDim.x) / LPR;
  float4 acc = make_float4(0, 0, 0, 0);
  for (long base = group * U; base + U <= n_gather; base += n_groups * U) {
    float4 x[U];
#pragma unroll
    for (int j = 0; j < U; ++j) x[j] = __ldg(X + (long)__ldg(idx + base + j) * LPR + lig);
#pragma unroll
    for (int j = 0; j < U; ++j) { acc.x += x[j].x; acc.y += x[j].y; acc.z += x[j].z; acc.w += x[j].w; }
  }
  if (acc.x + acc.y + acc.z + acc.w == 123.456f) out[0] = acc.x;   // keep the loads alive
}

template <int LPR, int U>
static void run(const char* name, int n_rows, long n_gather, int blocks_per_sm) {
  const int rowb = LPR * 16;
  float4* X; int* idx; float* out;
  cudaMalloc(&X, (size_t)n_rows * rowb); cudaMalloc(&idx, n_gather * 4); cudaMalloc(&out, 4);
  cudaMemset(X, 0, (size_t)n_rows * rowb);
  std::vector<int> h(n_gather);
  srand(1);
  for (long i = 0; i < n_gather; ++i) h[i] = (int)(((long)rand() * 32768 + rand()) % n_rows);
  cudaMemcpy(idx, h.data(), n_gather * 4, cudaMemcpyHostToDevice);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9;
  for (int rep = 0; rep < 12; ++rep) {
    cudaEventRecord(a);
    probe<LPR, U><<<sms * blocks_per_sm, 256>>>(X, idx, n_gather, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (rep >= 2 && ms < best) best = ms;
  }
  printf("%-34s rows=%8d table=%7.1f MB gathers=%9ld U=%2d occ=%d : %8.3f ms  %7.2f TB/s  %7.2f G rows/s\n", name, n_rows,
         n_rows * (double)rowb / 1e6, n_gather, U, blocks_per_sm, best, n_gather * (double)rowb / best / 1e9,
         n_gather / best / 1e6);
  cudaFree(X); cudaFree(idx); cudaFree(out);
}

int main() {
  // Amazon-Book shape: N = 144 242 rows of 256 B (d=64 fp32), nnz = 5 968 216 gathers per layer
  run<16, 4>("amazon d=64  (256 B rows)", 144242, 5968216, 4);
  run<16, 8>("amazon d=64  (256 B rows)", 144242, 5968216, 4);
  run<16, 8>("amazon d=64  (256 B rows)", 144242, 5968216, 8);
  run<16, 4>("amazon d=64  (256 B rows)", 144242, 5968216, 8);
  run<16, 16>("amazon d=64  (256 B rows)", 144242, 5968216, 4);
  run<16, 8>("gowalla d=64 (256 B rows)", 70839, 2054740, 8);
  // d=128 rows (512 B): synth-10m (120 K rows, 61 MB) and a table far beyond L2 (1.2 M rows, 614 MB)
  run<32, 8>("synth-10m d=128 (512 B rows)", 120000, 20000000, 8);
  run<32, 8>("synth-100m d=128, table >> L2", 1200000, 40000000, 8);
  run<32, 8>("12 M rows d=128, 6.1 GB table", 12000000, 40000000, 8);
  return 0;
}
