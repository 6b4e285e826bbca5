// Integer-throughput probe: the denominator of the ALU roofline (SURVEY.md 8d asks for the MEASURED
// min/add rate of the box instead of SMs x lanes x clock).  Three instruction mixes, each the inner
// operation of a path step (path_common.cuh) run as 8 independent dependency chains per thread:
//   kind 0  int32 add + min                 (ptxas fuses it into one 32-bit VIADDMNMX: 1 cell x 2 operations)
//   kind 1  packed u16x2 add, then min      (VIADD + VIMNMX.U16x2: 2 instructions for 2 cells x 2 operations)
//   kind 2  packed u16x2 fused add-min      (VIADDMNMX.U16x2, one DPX instruction = 2 cells x 2 operations)
#include "v3d_internal.h"

namespace {

constexpr int PROBE_CHAINS = 8;
constexpr int PROBE_THREADS = 256;

template <int KIND>
__global__ void __launch_bounds__(PROBE_THREADS) k_probe_alu(uint32_t* out, int iters, uint32_t p, uint32_t q)
{
    uint32_t a[PROBE_CHAINS];
#pragma unroll
    for (int k = 0; k < PROBE_CHAINS; k++) a[k] = (threadIdx.x * 0x00010001u + k * 0x00030005u) & 0x3fff3fffu;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int k = 0; k < PROBE_CHAINS; k++) {
                if (KIND == 0) a[k] = (uint32_t)min((int)(a[k] + p), (int)q);
                else if (KIND == 1) a[k] = __vminu2(a[k] + p, q);
                else a[k] = __viaddmin_u16x2(a[k], p, q);
            }
        }
        a[0] ^= (uint32_t)i;      // loop-dependent input; 1 instruction in 33, not counted
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < PROBE_CHAINS; k++) s += a[k];
    if (s == 0x12345678u) out[0] = s;      // practically never; defeats dead-code elimination
}

}  // namespace

extern "C" int v3d_probe_int_throughput(int device, int kind, double* lane_instr_per_s, double* algorithmic_ops_per_s)
{
    if (!lane_instr_per_s || !algorithmic_ops_per_s) return v3d_fail(V3D_EINVAL, "null argument");
    if (kind < 0 || kind > 2) return v3d_fail(V3D_EINVAL, "probe kind %d unknown (0..2)", kind);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return v3d_fail(V3D_ECUDA, "no CUDA device: libv3d has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return v3d_fail(V3D_EINVAL, "device %d out of range (%d devices)", device, ndev);
    V3D_CUDA(cudaSetDevice(device));
    int sms = 0;
    V3D_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    uint32_t* out = nullptr;
    V3D_CUDA(cudaMalloc(&out, sizeof(uint32_t)));
    cudaStream_t st;
    cudaEvent_t e0, e1;
    V3D_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    V3D_CUDA(cudaEventCreate(&e0));
    V3D_CUDA(cudaEventCreate(&e1));
    const int grid = sms * 8, iters = 256;                 // 8 CTAs x 8 warps per SM = full occupancy
    const uint32_t p = 0x00030003u, q = 0x3fff3fffu;
    int rc = V3D_OK;
    float best_ms = 0.f;
    for (int rep = 0; rep < 4 && rc == V3D_OK; rep++) {    // first pass = warm-up
        cudaEventRecord(e0, st);
        if (kind == 0) k_probe_alu<0><<<grid, PROBE_THREADS, 0, st>>>(out, iters, p, q);
        else if (kind == 1) k_probe_alu<1><<<grid, PROBE_THREADS, 0, st>>>(out, iters, p, q);
        else k_probe_alu<2><<<grid, PROBE_THREADS, 0, st>>>(out, iters, p, q);
        rc = v3d_cuda_check(cudaGetLastError(), "probe launch");
        cudaEventRecord(e1, st);
        if (rc == V3D_OK) rc = v3d_cuda_check(cudaEventSynchronize(e1), "probe run");
        float ms = 0.f;
        if (rc == V3D_OK) cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && (best_ms == 0.f || ms < best_ms)) best_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    cudaFree(out);
    if (rc != V3D_OK) return rc;
    if (!(best_ms > 0.f)) return v3d_fail(V3D_ECUDA, "probe measured no time");
    // per thread: iters x 4 x PROBE_CHAINS chain steps; a step is 1 lane-instruction for kinds 0 and 2 (fused
    // add-min, checked in the SASS), 2 for kind 1; algorithmic operations per step: add + min on 1 cell (kind 0)
    // or 2 cells (kinds 1, 2)
    const double steps = (double)grid * PROBE_THREADS * iters * 4.0 * PROBE_CHAINS;
    const double sec = best_ms * 1e-3;
    *lane_instr_per_s = steps * (kind == 1 ? 2.0 : 1.0) / sec;
    *algorithmic_ops_per_s = steps * (kind == 0 ? 2.0 : 4.0) / sec;
    return V3D_OK;
}
