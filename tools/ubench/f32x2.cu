// Micro-benchmark: issue/throughput of scalar FMUL+FADD vs packed FMUL2 + FFMA2(x, ONE, y) on sm_100a,
// and a bit-exactness check of the packed forms against __fmul_rn/__fadd_rn/__fsub_rn.
// (ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even with --fmad false, so the packed
// add is written as fma(a, ONE, b) with ONE opaque to the compiler.)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, u64 one, int iters, float seed) {
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 0.001f + i; b[i] = 1.0f + i * 1e-7f; }
  if (MODE == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] = __fadd_rn(__fmul_rn(a[i], b[i]), b[i]); } } }
  else if (MODE == 1) {
    u64 A[4], B[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { A[i] = pk(a[2*i], a[2*i+1]); B[i] = pk(b[2*i], b[2*i+1]); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { A[i] = fma2(mul2(A[i], B[i]), one, B[i]); } }
#pragma unroll
    for (int i = 0; i < 4; ++i) { upk(A[i], a[2*i], a[2*i+1]); } }
  else if (MODE == 2) {   // scalar FP + as many integer ops
    int k[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { k[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] = __fadd_rn(__fmul_rn(a[i], b[i]), b[i]); k[i] = (k[i] ^ (k[i] >> 3)) + it; } }
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] += k[i]; } }
  else {   // packed FP + the same integer ops
    u64 A[4], B[4];
    int k[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { k[i] = threadIdx.x + i; }
#pragma unroll
    for (int i = 0; i < 4; ++i) { A[i] = pk(a[2*i], a[2*i+1]); B[i] = pk(b[2*i], b[2*i+1]); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { A[i] = fma2(mul2(A[i], B[i]), one, B[i]); }
#pragma unroll
      for (int i = 0; i < 8; ++i) { k[i] = (k[i] ^ (k[i] >> 3)) + it; } }
#pragma unroll
    for (int i = 0; i < 4; ++i) { upk(A[i], a[2*i], a[2*i+1]); }
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] += k[i]; } }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += a[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s; }

__global__ void exact(const uint32_t* x, const uint32_t* y, int n, u64 one, u64 negone, unsigned long long* bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i * 2 + 1 >= n) return;
  float a0 = __uint_as_float(x[2*i]), a1 = __uint_as_float(x[2*i+1]);
  float b0 = __uint_as_float(y[2*i]), b1 = __uint_as_float(y[2*i+1]);
  float r0, r1;
  unsigned long long nb = 0;
  upk(mul2(pk(a0, a1), pk(b0, b1)), r0, r1);
  nb += (__float_as_uint(r0) != __float_as_uint(__fmul_rn(a0, b0))) + (__float_as_uint(r1) != __float_as_uint(__fmul_rn(a1, b1)));
  upk(fma2(pk(a0, a1), one, pk(b0, b1)), r0, r1);
  nb += (__float_as_uint(r0) != __float_as_uint(__fadd_rn(a0, b0))) + (__float_as_uint(r1) != __float_as_uint(__fadd_rn(a1, b1)));
  upk(fma2(pk(b0, b1), negone, pk(a0, a1)), r0, r1);
  nb += (__float_as_uint(r0) != __float_as_uint(__fsub_rn(a0, b0))) + (__float_as_uint(r1) != __float_as_uint(__fsub_rn(a1, b1)));
  // mul then add, the pattern ptxas would contract
  upk(fma2(mul2(pk(a0, a1), pk(b0, b1)), one, pk(a0, a1)), r0, r1);
  nb += (__float_as_uint(r0) != __float_as_uint(__fadd_rn(__fmul_rn(a0, b0), a0))) + (__float_as_uint(r1) != __float_as_uint(__fadd_rn(__fmul_rn(a1, b1), a1)));
  if (nb) atomicAdd(bad, nb); }

int main() {
  float one2[2] = {1.0f, 1.0f}, neg2[2] = {-1.0f, -1.0f};
  u64 one, negone; memcpy(&one, one2, 8); memcpy(&negone, neg2, 8);
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) bench<0><<<148 * 8, 256>>>(out, one, iters, 1.0f);
      if (mode == 1) bench<1><<<148 * 8, 256>>>(out, one, iters, 1.0f);
      if (mode == 2) bench<2><<<148 * 8, 256>>>(out, one, iters, 1.0f);
      if (mode == 3) bench<3><<<148 * 8, 256>>>(out, one, iters, 1.0f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double flops = 148.0 * 8 * 256 * iters * 16.0;   // 8 mul + 8 add per iteration per thread
      if (rep) printf("mode %d: %.3f ms  %.2f TFLOP/s (mul+add, unfused)\n", mode, ms, flops / ms * 1e-9); } }
  // exactness on random bit patterns (all exponents, denormals, infs; NaNs excluded)
  const int n = 1 << 24;
  uint32_t* hx = (uint32_t*)malloc(n * 4); uint32_t* hy = (uint32_t*)malloc(n * 4);
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); };
  for (int i = 0; i < n; ++i) {
    uint32_t a = rnd(), b = rnd();
    if (i & 1) { b = (b & 0x807fffffu) | (((a >> 23) & 0xff) + (rnd() % 5) - 2) << 23; }   // nearby exponents: cancellation
    if (((a >> 23) & 0xff) == 0xff && (a & 0x7fffff)) a &= 0xff800000u;
    if (((b >> 23) & 0xff) == 0xff && (b & 0x7fffff)) b &= 0xff800000u;
    hx[i] = a; hy[i] = b; }
  uint32_t *dx, *dy; unsigned long long* dbad; cudaMalloc(&dx, n * 4); cudaMalloc(&dy, n * 4); cudaMalloc(&dbad, 8);
  cudaMemcpy(dx, hx, n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dy, hy, n * 4, cudaMemcpyHostToDevice); cudaMemset(dbad, 0, 8);
  exact<<<(n / 2 + 255) / 256, 256>>>(dx, dy, n, one, negone, dbad);
  unsigned long long bad = 0; cudaMemcpy(&bad, dbad, 8, cudaMemcpyDeviceToHost);
  printf("exactness: %llu mismatches over %d pairs x 4 ops (inf-inf / 0*inf NaN results compare by bits)\n", bad, n / 2);
  printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0; }
