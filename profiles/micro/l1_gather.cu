// l1_gather.cu -- microbenchmark: what does a per-lane 32-byte gather cost in the L1 data pipe, as a function of how the 32
// lanes of one instruction are spread over 128-byte lines?  (Decides the design of the cutoff pair kernel: the Verlet-list
// kernel of csrc/nbx_cells.cu measured ~27 LSU wavefronts per warp gather.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1_gather l1_gather.cu && ./l1_gather
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ double4 ld256(const double4 *p)
{
    double4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}

// idx[step][lane] patterns are generated on the host: tab[steps][32] record indices (same for every warp, plus a per-warp offset)
template <int MODE> // 0: LDG.256 of the record, 1: LDG.128 (first half), 2: shared memory 3 x LDS.64 (SoA), 3: smem AoS LDS.128+LDS.64
__global__ void __launch_bounds__(128) gather_kernel(const double4 *__restrict__ rec, const int *__restrict__ tab, int steps, int T,
                                                     int reps, double *__restrict__ out, int hashed)
{
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int woff = ((blockIdx.x * 4 + warp) * 37) % T;
    if (MODE >= 2) {
        for (int i = threadIdx.x; i < T; i += 128) {
            const double4 r = rec[i];
            if (MODE == 2) { sm[i] = r.x; sm[T + i] = r.y; sm[2 * T + i] = r.z; }
            else { reinterpret_cast<double4 *>(sm)[i] = r; }
        }
        __syncthreads();
    }
    double acc = 0.0;
    for (int r = 0; r < reps; ++r) {
        for (int s = 0; s + 4 <= steps; s += 4) {
            int i0 = tab[s * 32 + lane] + woff, i1 = tab[(s + 1) * 32 + lane] + woff, i2 = tab[(s + 2) * 32 + lane] + woff,
                i3 = tab[(s + 3) * 32 + lane] + woff;
            i0 = i0 >= T ? i0 - T : i0; i1 = i1 >= T ? i1 - T : i1; i2 = i2 >= T ? i2 - T : i2; i3 = i3 >= T ? i3 - T : i3;
            if (hashed) { // a different random record per warp, step and lane (position = lane mod 4): L1 misses when T is large
                unsigned h = (unsigned)((blockIdx.x * 4 + warp) * 64 + s) * 2654435761u + (unsigned)lane * 40503u + (unsigned)r * 97u;
                auto mix = [&](unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; };
                const unsigned q = (unsigned)T / 4u;
                i0 = (int)(mix(h) % q) * 4 + (lane & 3); i1 = (int)(mix(h + 1u) % q) * 4 + (lane & 3);
                i2 = (int)(mix(h + 2u) % q) * 4 + (lane & 3); i3 = (int)(mix(h + 3u) % q) * 4 + (lane & 3);
            }
            if (MODE == 0) {
                const double4 a = ld256(rec + i0), b = ld256(rec + i1), c = ld256(rec + i2), d = ld256(rec + i3);
                acc += a.x + b.y + c.z + d.x + a.w;
            } else if (MODE == 1) {
                const double2 a = __ldg(reinterpret_cast<const double2 *>(rec + i0)), b = __ldg(reinterpret_cast<const double2 *>(rec + i1)),
                              c = __ldg(reinterpret_cast<const double2 *>(rec + i2)), d = __ldg(reinterpret_cast<const double2 *>(rec + i3));
                acc += a.x + b.y + c.x + d.y;
            } else if (MODE == 2) {
                acc += sm[i0] + sm[T + i0] + sm[2 * T + i0] + sm[i1] + sm[T + i1] + sm[2 * T + i1] + sm[i2] + sm[T + i2] + sm[2 * T + i2] +
                       sm[i3] + sm[T + i3] + sm[2 * T + i3];
            } else {
                const double4 *q = reinterpret_cast<const double4 *>(sm);
                const double2 a = *reinterpret_cast<const double2 *>(q + i0), b = *reinterpret_cast<const double2 *>(q + i1),
                              c = *reinterpret_cast<const double2 *>(q + i2), d = *reinterpret_cast<const double2 *>(q + i3);
                acc += a.x + a.y + q[i0].z + b.x + b.y + q[i1].z + c.x + c.y + q[i2].z + d.x + d.y + q[i3].z;
            }
        }
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}

static unsigned rng_state = 12345u;
static unsigned rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

int main(int argc, char **argv)
{
    const int T = argc > 1 ? atoi(argv[1]) : 1536;      // records (48 KB): L1-resident, and the size of a staged tile in shared memory
    const int steps = 64, reps = T > 1536 ? 40 : 200;
    int sms = 148;
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0); sms = prop.multiProcessorCount;
    double4 *rec; int *tab; double *out;
    cudaMalloc(&rec, sizeof(double4) * T); cudaMalloc(&tab, sizeof(int) * steps * 32);
    const int blocks = sms * 8;
    cudaMalloc(&out, sizeof(double) * blocks * 128);
    std::vector<double4> h(T);
    for (int i = 0; i < T; ++i) h[i] = make_double4(i, 2 * i, 3 * i, 1.0);
    cudaMemcpy(rec, h.data(), sizeof(double4) * T, cudaMemcpyHostToDevice);
    if (T <= 1536) {
        cudaFuncSetAttribute(gather_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * T * 8);
        cudaFuncSetAttribute(gather_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * T * 8);
    }
    struct Pat { const char *name; int kind; };
    const Pat pats[] = {{"32 lines x 1 sector (stride 4 records)", 0}, {"8 lines x 4 sectors (consecutive records)", 1},
                        {"16 lines x 2 sectors", 2}, {"broadcast: 1 record", 3}, {"8 lines x 4 sectors, lanes permuted", 4},
                        {"16 records, each read by 2 lanes, 16 lines", 5}, {"4 records x 8 lanes, 1 line", 6},
                        {"random over the table", 7}, {"random within a 64-record window (16 lines)", 8},
                        {"random within a 32-record window (8 lines)", 9}, {"sorted random within 64-record window", 10},
                        {"random in 64-window, lane l -> sector position l%4", 11}, {"random over table, lane l -> sector position l%4", 12},
                        {"random in 64-window, 3 of 4 steps position l%4", 13},
                        {"hashed per warp/step/lane over the table, position l%4", 14}};
    std::vector<int> ht(steps * 32);
    for (const Pat &p : pats) {
        for (int s = 0; s < steps; ++s) {
            const int base = (s * 211) % T;
            int perm[32];
            for (int l = 0; l < 32; ++l) perm[l] = l;
            for (int l = 31; l > 0; --l) { int j = rnd() % (l + 1); int t = perm[l]; perm[l] = perm[j]; perm[j] = t; }
            int tmp[32];
            for (int l = 0; l < 32; ++l) {
                int idx = 0;
                switch (p.kind) {
                case 0: idx = base / 4 * 4 + l * 4; break;
                case 1: idx = base / 4 * 4 + l; break;
                case 2: idx = base / 4 * 4 + (l / 2) * 4 + (l % 2); break;
                case 3: idx = base; break;
                case 4: idx = base / 4 * 4 + perm[l]; break;
                case 5: idx = base / 4 * 4 + (l / 2) * 4; break;
                case 6: idx = base / 4 * 4 + (l / 8); break;
                case 7: idx = rnd() % T; break;
                case 8: idx = base / 4 * 4 + rnd() % 64; break;
                case 9: idx = base / 4 * 4 + rnd() % 32; break;
                case 10: idx = base / 4 * 4 + rnd() % 64; break;
                case 11: idx = base / 4 * 4 + 4 * (rnd() % 16) + (l % 4); break;
                case 12: idx = 4 * (rnd() % (T / 4)) + (l % 4); break;
                case 13: idx = (s % 4 == 3) ? base / 4 * 4 + rnd() % 64 : base / 4 * 4 + 4 * (rnd() % 16) + (l % 4); break;
                }
                tmp[l] = idx % T;
            }
            if (p.kind == 10) { for (int a = 0; a < 32; ++a) for (int b = a + 1; b < 32; ++b) if (tmp[b] < tmp[a]) { int t = tmp[a]; tmp[a] = tmp[b]; tmp[b] = t; } }
            for (int l = 0; l < 32; ++l) ht[s * 32 + l] = tmp[l];
        }
        cudaMemcpy(tab, ht.data(), sizeof(int) * steps * 32, cudaMemcpyHostToDevice);
        printf("%-52s", p.name);
        for (int mode = 0; mode < (T > 1536 ? 2 : 4); ++mode) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            const size_t smem = mode == 2 ? 3 * T * 8 : (mode == 3 ? 4 * T * 8 : 0);
            for (int it = 0; it < 2; ++it) {
                if (it == 1) cudaEventRecord(e0);
                if (mode == 0) gather_kernel<0><<<blocks, 128, smem>>>(rec, tab, steps, T, reps, out, p.kind == 14);
                else if (mode == 1) gather_kernel<1><<<blocks, 128, smem>>>(rec, tab, steps, T, reps, out, p.kind == 14);
                else if (mode == 2) gather_kernel<2><<<blocks, 128, smem>>>(rec, tab, steps, T, reps, out, p.kind == 14);
                else gather_kernel<3><<<blocks, 128, smem>>>(rec, tab, steps, T, reps, out, p.kind == 14);
                if (it == 1) cudaEventRecord(e1);
            }
            cudaDeviceSynchronize();
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            int clk_khz = prop.clockRate; // kHz
            // warp-gathers per SM: blocks/sms * 4 warps * steps * reps
            const double wg = (double)blocks / sms * 4.0 * steps * reps;
            const double cyc = ms * 1e-3 * (double)clk_khz * 1e3 / wg;
            printf("  %s %6.2f", mode == 0 ? "LDG256" : mode == 1 ? "LDG128" : mode == 2 ? "LDS64x3" : "LDSaos", cyc);
        }
        printf("   (SM cycles per warp gather at %d MHz nominal)\n", prop.clockRate / 1000);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    }
    return 0;
}
