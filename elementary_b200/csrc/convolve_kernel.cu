// convolve_kernel.cu — K3: uniformly partitioned FFT convolution, hand-written shared-memory FFT + complex MAC.
// No cuFFT, no tensor cores: this is a streaming frequency-domain delay line, bound by the HBM reads of the input
// spectra (see convolve.h for the algorithmic bytes).
//
// Persistent CTAs (2 per SM; 8 consumer warps + 1 TMA producer warp) walk the channel pairs; a pair shares one pass over
// the IR spectra so that every IR spectrum value fetched from L2 is reused.  Per pair and call:
//   1. append the new samples to the partition's input buffer, zero-pad to 1024 (FFTConvolver.cpp:157-164)
//   2. real FFT 1024 = complex Stockham radix-2 FFT 512 in shared memory + split post-pass  (replaces OouraFFT::fft,
//      AudioFFT.cpp:132-155; float arithmetic instead of the reference's double)
//   3. if the input buffer was empty: Ypre = sum_{i>=1} H_i * X_{cur+i}   (FFTConvolver.cpp:168-177,
//      ComplexMultiplyAccumulate Utilities.cpp:66-117)
//   4. Y = Ypre + X_cur * H_0 (:178-179); inverse real FFT (:182); out = y[fill..] + overlap[fill..] (:185)
//   5. when the partition is complete: overlap = y[512..1024) (:194), the host rotates `cur` (:200)
#include "convolve.h"
#include <cstdio>

namespace eb {

namespace {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// ---- TMA (1-D bulk async copy) + mbarrier primitives: the frequency-domain delay line is streamed HBM -> shared memory
// by cp.async.bulk (SASS: UBLKCP) through a multi-stage ring guarded by full/empty mbarriers and fed by a dedicated
// producer warp, so the MAC loop never waits on a per-thread global load and the stream runs through the FFT phases.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int CH = CONV_CH_PER_CTA;
#ifndef EB_CONV_STAGES
#define EB_CONV_STAGES 2   /* ring depth; 2 x 12 KB + 20 KB work area = 44 KB per CTA -> 4 persistent CTAs per SM (A/B: profiles/r01_n_k3_persistent_ab.txt) */
#endif
constexpr int STAGES = EB_CONV_STAGES;                       // delay-line pipeline depth: STAGES x (CH + 1) rows of 4 KB in flight per CTA
constexpr uint32_t ROW_BYTES = CONV_PACKED_BINS * sizeof(float2);
constexpr int NB = CONV_PACKED_BINS;   // 512: bin 0 carries (Re X[0], Re X[512]) — both are purely real
constexpr int N2 = 512;   // complex FFT length
constexpr int CONSUMERS = 256;         // 8 consumer warps (FFT + MAC); warp 8 is the TMA producer
constexpr int CONV_THREADS = CONSUMERS + 32;
#ifndef EB_CONV_CTAS
#define EB_CONV_CTAS 4
#endif
constexpr int CONV_CTAS_PER_SM = EB_CONV_CTAS;

// barrier among the consumer warps only (the producer warp never joins): named barrier 1
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 9 Stockham radix-2 stages over CH independent 512-point transforms; 256 threads = one butterfly per thread per
// channel per stage.  The result lands in `b`.
template <bool INVERSE>
__device__ __forceinline__ void fft512(float2 (*a)[N2], float2 (*b)[N2], const float2* tw, int tid) {
    float2 (*src)[N2] = a;
    float2 (*dst)[N2] = b;
#pragma unroll 1
    for (int ns = 1; ns < N2; ns <<= 1) {
        const int k = tid & (ns - 1);
        float2 w = tw[k * (N2 / ns)];
        if (INVERSE) w.y = -w.y;
        const int j0 = ((tid - k) << 1) + k;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const float2 u = src[c][tid];
            const float2 v = cmul(w, src[c][tid + N2 / 2]);
            dst[c][j0] = cadd(u, v);
            dst[c][j0 + ns] = csub(u, v);
        }
        consumer_sync();
        float2 (*t)[N2] = src; src = dst; dst = t;
    }
}

} // namespace

// Persistent, warp-specialised kernel: CONV_CTAS_PER_SM CTAs per SM, each walking the channel pairs
// u = blockIdx.x, blockIdx.x + gridDim.x, ...  Inside a CTA
//   * the producer warp streams the frequency-domain delay line (and the matching IR spectra, L2-resident) HBM -> shared
//     memory with cp.async.bulk through a STAGES-deep ring of full/empty mbarriers, running ahead of the consumers —
//     across the consumers' FFT phases and into the next channel pair — so the HBM stream never waits for math;
//   * the 8 consumer warps do the forward FFT of the new block, the complex MAC over the ring, the inverse FFT and the
//     overlap-add.  The MAC over the older partitions does not depend on the new block at all, which is what lets the
//     producer start a pair's stream before its FFT has even begun.
__global__ void __launch_bounds__(CONV_THREADS, CONV_CTAS_PER_SM) convolve_chunk_kernel(
    const float* __restrict__ in, float* __restrict__ out, int stride, int offset, int n, int fill, int cur, int S, int nv,
    const float2* __restrict__ H, float2* __restrict__ fdl, float2* __restrict__ ypre,
    float* __restrict__ overlap, float* __restrict__ inbuf, const float2* __restrict__ twg, int smCount) {
    extern __shared__ __align__(128) unsigned char smemRaw[];
    float2 (*stX)[CH][N2] = reinterpret_cast<float2 (*)[CH][N2]>(smemRaw);                                   // [STAGES][CH][512]
    float2 (*stH)[N2] = reinterpret_cast<float2 (*)[N2]>(smemRaw + (size_t) STAGES * CH * ROW_BYTES);        // [STAGES][512]
    float2 (*A)[N2] = reinterpret_cast<float2 (*)[N2]>(smemRaw + (size_t) STAGES * (CH + 1) * ROW_BYTES);    // [CH][512]
    float2 (*B)[N2] = A + CH;                                                                                 // [CH][512]
    float2* tw = reinterpret_cast<float2*>(B + CH);                                                           // [512]
    uint64_t* full = reinterpret_cast<uint64_t*>(tw + N2);                                                    // [STAGES]
    uint64_t* empty = full + STAGES;                                                                          // [STAGES]

    const int tid = threadIdx.x;
    const int numUnits = (nv + CH - 1) / CH;

    if (tid == 0) {
        for (int sidx = 0; sidx < STAGES; ++sidx) { mbar_init(&full[sidx], 1); mbar_init(&empty[sidx], CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    if (tid < CONSUMERS) { tw[tid] = twg[tid]; tw[tid + 256] = twg[tid + 256]; }
    __syncthreads();

    if (tid >= CONSUMERS) {
        // ---------------- producer warp: one elected lane issues every bulk copy ----------------
        if (tid == CONSUMERS && fill == 0) {
            uint32_t it = 0;                                         // stage uses so far (shared numbering with the consumers)
            for (int u = blockIdx.x; u < numUnits; u += gridDim.x) {
                const int ch0 = u * CH;
                for (int i = 1; i < S; ++i, ++it) {
                    const int stg = it % STAGES;
                    mbar_wait(&empty[stg], ((it / STAGES) & 1) ^ 1);   // the consumers are done with the stage's previous contents
                    int slotIdx = cur + i;
                    if (slotIdx >= S) slotIdx -= S;
                    mbar_expect_tx(&full[stg], (CH + 1) * ROW_BYTES);
#pragma unroll
                    for (int c = 0; c < CH; ++c)
                        tma_load_1d(&stX[stg][c][0], fdl + ((size_t) min(ch0 + c, nv - 1) * S + slotIdx) * NB, ROW_BYTES, &full[stg]);
                    tma_load_1d(&stH[stg][0], H + (size_t) i * NB, ROW_BYTES, &full[stg]);
                }
            }
        }
        return;
    }

    // ---------------- consumer warps ----------------
    // The MAC over the older partitions does not depend on the block that just arrived, so a CTA may run it before or
    // after the forward FFT.  The two CTAs that share an SM use opposite orders: while one streams the delay line the
    // other does its transforms, and the HBM stream of the SM never pauses for an FFT phase.
    const bool macFirst = ((blockIdx.x / smCount) & 1) != 0;   // CTAs b, b + smCount, b + 2*smCount, ... land on the same SM
    const bool wholeBlock = (fill == 0) && (n == CONV_BLOCK) && (((stride | offset) & 1) == 0);   // no staging through inbuf needed
    const float scale = 1.0f / 512.0f;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < numUnits; u += gridDim.x) {
        const int ch0 = u * CH;
        float2 acc[2][CH];

        // 3. frequency-domain delay line: acc[b] = sum_{i>=1} H_i[b] * X_{cur+i}[b], once per block (fill == 0).
        // Bin 0 is the packed pair of real bins and multiplies component-wise.  Each thread owns bins tid and tid+256 of
        // CH channels and reads them from the stage the producer filled.
        auto macPhase = [&]() {
            if (fill == 0) {
#pragma unroll
                for (int c = 0; c < CH; ++c) { acc[0][c] = make_float2(0.0f, 0.0f); acc[1][c] = make_float2(0.0f, 0.0f); }
                for (int i = 1; i < S; ++i, ++it) {
                    const int stg = it % STAGES;
                    mbar_wait(&full[stg], (it / STAGES) & 1);                // bytes of partition i have landed
                    const float2 ha = stH[stg][tid], hb = stH[stg][tid + 256];
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const float2 xa = stX[stg][c][tid];
                        const float2 xb = stX[stg][c][tid + 256];
                        acc[0][c] = (tid == 0) ? make_float2(acc[0][c].x + ha.x * xa.x, acc[0][c].y + ha.y * xa.y) : cadd(acc[0][c], cmul(ha, xa));
                        acc[1][c] = cadd(acc[1][c], cmul(hb, xb));
                    }
                    __syncwarp();
                    if ((tid & 31) == 0) mbar_arrive(&empty[stg]);           // this warp is done with the stage
                }
#pragma unroll
                for (int c = 0; c < CH; ++c) if (ch0 + c < nv) {
                    ypre[(size_t) (ch0 + c) * NB + tid] = acc[0][c];
                    ypre[(size_t) (ch0 + c) * NB + tid + 256] = acc[1][c];
                }
            } else {
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const size_t o = (size_t) min(ch0 + c, nv - 1) * NB + tid;
                    acc[0][c] = ypre[o]; acc[1][c] = ypre[o + 256];
                }
            }
        };

        // 1./2. the new samples -> 512 complex points (even, odd), zero-padded -> FFT512 -> bins of the real FFT (A, and the
        // delay-line slot `cur`)
        auto forwardPhase = [&]() {
            if (wholeBlock) {
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int ch = ch0 + c;
                    float2 z = make_float2(0.0f, 0.0f);
                    if (ch < nv) z = __ldg(reinterpret_cast<const float2*>(in + (size_t) ch * stride + offset) + tid);
                    A[c][tid] = z;
                    A[c][tid + 256] = make_float2(0.0f, 0.0f);
                }
            } else {
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int ch = ch0 + c;
                    if (ch < nv) {
                        float* ib = inbuf + (size_t) ch * CONV_BLOCK;
                        for (int i = tid; i < n; i += CONSUMERS) ib[fill + i] = in[(size_t) ch * stride + offset + i];
                    }
                }
                consumer_sync();
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int ch = ch0 + c;
                    float2 z = make_float2(0.0f, 0.0f);
                    if (ch < nv) z = reinterpret_cast<const float2*>(inbuf + (size_t) ch * CONV_BLOCK)[tid];
                    A[c][tid] = z;
                    A[c][tid + 256] = make_float2(0.0f, 0.0f);
                }
            }
            consumer_sync();
            fft512<false>(A, B, tw, tid);      // Z = FFT512(z) lands in B
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const int ch = ch0 + c;
                float2* slot = fdl + ((size_t) ch * S + cur) * NB;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k = tid + h * 256;
                    const float2 zk = B[c][k];
                    const float2 zn = cconj(B[c][(N2 - k) & (N2 - 1)]);
                    const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y + zn.y));
                    const float2 d = csub(zk, zn);
                    const float2 o = make_float2(0.5f * d.y, -0.5f * d.x);      // -0.5i * (zk - zn)
                    float2 x = cadd(e, cmul(tw[k], o));
                    if (k == 0) x = make_float2(e.x + o.x, e.x - o.x);          // packed: (X[0], X[512]) = (E0 + O0, E0 - O0), both real
                    A[c][k] = x;
                    if (ch < nv) slot[k] = x;
                }
            }
            consumer_sync();
        };

#ifdef EB_CONV_TIMING
        unsigned long long tq0, tq1, tq2, tq3;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tq0));
        if (macFirst) { macPhase(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tq1)); forwardPhase(); }
        else { forwardPhase(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tq1)); macPhase(); }
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tq2));
#else
        if (macFirst) { macPhase(); forwardPhase(); } else { forwardPhase(); macPhase(); }
#endif

        // overlap of the previous partition for the samples of this chunk: fetched now, used after the inverse FFT
        float ovv[CH][2];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int ch = min(ch0 + c, nv - 1);
            const float* ov = overlap + (size_t) ch * CONV_BLOCK;
            ovv[c][0] = (tid < n) ? ov[fill + tid] : 0.0f;
            ovv[c][1] = (tid + 256 < n) ? ov[fill + tid + 256] : 0.0f;
        }

        // 4. Y = acc + X_cur * H_0 (FFTConvolver.cpp:178-179)
        {
            const float2 h0a = __ldg(H + tid), h0b = __ldg(H + tid + 256);
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const float2 xa = A[c][tid], xb = A[c][tid + 256];
                B[c][tid] = (tid == 0) ? make_float2(acc[0][c].x + xa.x * h0a.x, acc[0][c].y + xa.y * h0a.y) : cadd(acc[0][c], cmul(xa, h0a));
                B[c][tid + 256] = cadd(acc[1][c], cmul(xb, h0b));
            }
        }
        consumer_sync();

        // inverse split: Zi[k] = E[k] + i*O[k] from Y[k], conj(Y[512-k])  (reads B, writes A)
#pragma unroll
        for (int c = 0; c < CH; ++c) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = tid + h * 256;
                const float2 yk = B[c][k];
                float2 z;
                if (k == 0) {                                                 // packed (Y0, Y512): E0 = (Y0+Y512)/2, O0 = (Y0-Y512)/2
                    z = make_float2(0.5f * (yk.x + yk.y), 0.5f * (yk.x - yk.y));
                } else {
                    const float2 yn = cconj(B[c][N2 - k]);
                    const float2 e = make_float2(0.5f * (yk.x + yn.x), 0.5f * (yk.y + yn.y));
                    const float2 d = make_float2(0.5f * (yk.x - yn.x), 0.5f * (yk.y - yn.y));
                    const float2 o = cmul(d, cconj(tw[k]));
                    z = make_float2(e.x - o.y, e.y + o.x);                    // e + i*o
                }
                A[c][k] = z;
            }
        }
        consumer_sync();
        fft512<true>(A, B, tw, tid);   // result in B: z[j] * 512

        // 5. overlap-add output for the samples of this chunk; save the second half when the partition is complete
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int ch = ch0 + c;
            if (ch >= nv) continue;
            const float* y = reinterpret_cast<const float*>(&B[c][0]);      // y[2j] = re z[j], y[2j+1] = im z[j]
            float* o = out + (size_t) ch * stride + offset;
            if (tid < n) o[tid] = y[fill + tid] * scale + ovv[c][0];
            if (tid + 256 < n) o[tid + 256] = y[fill + tid + 256] * scale + ovv[c][1];
        }
        if (fill + n == CONV_BLOCK) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const int ch = ch0 + c;
                if (ch >= nv) continue;
                const float* y = reinterpret_cast<const float*>(&B[c][0]);
                float* ov = overlap + (size_t) ch * CONV_BLOCK;
                float* ib = inbuf + (size_t) ch * CONV_BLOCK;
                for (int i = tid; i < CONV_BLOCK; i += CONSUMERS) { ov[i] = y[CONV_BLOCK + i] * scale; if (!wholeBlock) ib[i] = 0.0f; }
            }
        }
        consumer_sync();   // A/B are reused by the next channel pair
#ifdef EB_CONV_TIMING
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tq3));
        if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 1 || blockIdx.x == 148 || blockIdx.x == 295 || blockIdx.x == 200))
            printf("TIMING cta %d unit %d macFirst %d : first %llu ns, second %llu ns, inverse+out %llu ns, start %llu\n", blockIdx.x, u, (int) macFirst,
                   tq1 - tq0, tq2 - tq1, tq3 - tq2, tq0 % 1000000ull);
#endif
    }
}

cudaError_t convolver_process_chunk(ConvolverState& st, const float* in, float* out, int stride, int offset, int n, cudaStream_t stream) {
    if (st.planOnly) return cudaErrorNotSupported;
    if (st.partitions == 0) {   // empty (fully trimmed) IR: silence (FFTConvolver.cpp:149-153)
        return cudaMemset2DAsync(out + offset, sizeof(float) * stride, 0, sizeof(float) * n, st.nv, stream);
    }
    const int units = (st.nv + CONV_CH_PER_CTA - 1) / CONV_CH_PER_CTA;
    const size_t smem = (size_t) STAGES * (CH + 1) * ROW_BYTES + (size_t) 2 * CH * ROW_BYTES + ROW_BYTES + 2 * STAGES * sizeof(uint64_t);
    static int persistentCtas = 0, smCount = 148;
    if (!persistentCtas) {
        cudaError_t e = cudaFuncSetAttribute(convolve_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e != cudaSuccess) return e;
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        smCount = sms > 0 ? sms : 148;
        persistentCtas = smCount * CONV_CTAS_PER_SM;                   // a multiple of the SM count: one resident wave
    }
    const int grid = units < persistentCtas ? units : persistentCtas;
    convolve_chunk_kernel<<<grid, CONV_THREADS, smem, stream>>>(in, out, stride, offset, n, st.fill, st.cur, st.partitions, st.nv,
                                                             st.dH, st.dFdl, st.dYpre, st.dOverlap, st.dInBuf, st.dTw, smCount);
    st.fill += n;
    if (st.fill == CONV_BLOCK) {
        st.fill = 0;
        st.cur = (st.cur > 0) ? st.cur - 1 : st.partitions - 1;   // FFTConvolver.cpp:200
    }
    return cudaGetLastError();
}

} // namespace eb
