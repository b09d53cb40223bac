// tools/lds_probe.cu -- is the consumers' tap path limited by LDS instruction count or by shared-memory bytes?
// 15 warps per CTA (one CTA per SM), each lane gathers the 4-channel 2x2 bilinear footprint of 4 pixels per "plane":
//   mode 0: planar layout [row][ch][x], 16 LDS.32 per pixel (what the forward kernel does)
//   mode 1: texel-interleaved layout [row][x][ch], 4 LDS.128 per pixel (same bytes)
// scale = texel step per pixel (1.0: conflict-free rows; 1.2: 32 lanes span 38 texels)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int BW = 88, ROWS = 40;

template <int MODE>
__global__ void __launch_bounds__(480, 1) k(float* out, int planes, float scale) {
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < BW * ROWS * 4; i += blockDim.x) sm[i] = (float)(i & 1023) * 1e-3f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int p = 0; p < planes; ++p) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int x0 = (int)((lane + 32 * (q & 1)) * scale * 0.5f + (p & 3)), y0 = 2 * warp + (q >> 1) + (p & 1);
            if (MODE == 0) {
                const float* t = sm + (y0 * 4 * BW + x0);
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[c] += t[c * BW] + t[c * BW + 1] + t[(4 + c) * BW] + t[(4 + c) * BW + 1];
            } else {
                const float4* t = reinterpret_cast<const float4*>(sm) + (y0 * BW + x0);
                const float4 a = t[0], b = t[1], c2 = t[BW], d = t[BW + 1];
                acc[0] += a.x + b.x + c2.x + d.x; acc[1] += a.y + b.y + c2.y + d.y;
                acc[2] += a.z + b.z + c2.z + d.z; acc[3] += a.w + b.w + c2.w + d.w;
            }
        }
    }
    if (acc[0] + acc[1] + acc[2] + acc[3] == 12345.678f) out[0] = acc[0];
}

int main() {
    float* out; cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int planes = 20000; const size_t smem = BW * ROWS * 16;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (float scale : {1.9f, 2.4f}) {      // x0 uses scale*0.5: 0.95 and 1.2 texels per pixel
        for (int mode = 0; mode < 2; ++mode) {
            float ms;
            if (mode == 0) { k<0><<<148, 480, smem>>>(out, 100, scale); cudaEventRecord(e0); k<0><<<148, 480, smem>>>(out, planes, scale); }
            else { k<1><<<148, 480, smem>>>(out, 100, scale); cudaEventRecord(e0); k<1><<<148, 480, smem>>>(out, planes, scale); }
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            const double wp = (double)planes * 15 * 4;      // warp-planes per SM
            printf("scale %.2f mode %d (%s): %.3f ms -> %.1f SM-cycles per warp-plane (@1.965 GHz)  %s\n", scale * 0.5f, mode,
                   mode ? "4 x LDS.128 interleaved" : "16 x LDS.32 planar", ms, ms * 1e-3 * 1.965e9 / wp, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
