// convolve_kernel.cu — K3: uniformly partitioned FFT convolution, hand-written shared-memory FFT + complex MAC.
// No cuFFT, no tensor cores: this is a streaming frequency-domain delay line, bound by the HBM reads of the input
// spectra (see convolve.h for the algorithmic bytes).
//
// One CTA (256 threads) serves CONV_CH_PER_CTA channels so that every IR spectrum value fetched from L2 is reused.  Per call:
//   1. append the new samples to the partition's input buffer, zero-pad to 1024 (FFTConvolver.cpp:157-164)
//   2. real FFT 1024 = complex Stockham radix-2 FFT 512 in shared memory + split post-pass  (replaces OouraFFT::fft,
//      AudioFFT.cpp:132-155; float arithmetic instead of the reference's double)
//   3. if the input buffer was empty: Ypre = sum_{i>=1} H_i * X_{cur+i}   (FFTConvolver.cpp:168-177,
//      ComplexMultiplyAccumulate Utilities.cpp:66-117)
//   4. Y = Ypre + X_cur * H_0 (:178-179); inverse real FFT (:182); out = y[fill..] + overlap[fill..] (:185)
//   5. when the partition is complete: overlap = y[512..1024) (:194), the host rotates `cur` (:200)
#include "convolve.h"

namespace eb {

namespace {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// ---- TMA (1-D bulk async copy) + mbarrier primitives: the frequency-domain delay line is streamed HBM -> shared memory
// by cp.async.bulk (SASS: UBLKCP) through a multi-stage ring guarded by mbarriers, so the MAC loop never waits on a
// per-thread global load and the first stages are already in flight while the forward FFT runs.
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
#define EB_CONV_STAGES 2   /* measured (profiles/r01_k_k3_tma_stages.txt): 2 stages keep 5 CTAs per SM = one wave for 512 CTAs */
#endif
constexpr int STAGES = EB_CONV_STAGES;                       // delay-line pipeline depth: STAGES x (CH + 1) rows of 4 KB in flight per CTA
constexpr uint32_t ROW_BYTES = CONV_PACKED_BINS * sizeof(float2);
constexpr int NB = CONV_PACKED_BINS;   // 512: bin 0 carries (Re X[0], Re X[512]) — both are purely real
constexpr int N2 = 512;   // complex FFT length

// 9 Stockham radix-2 stages over CH independent 512-point transforms; 256 threads = one butterfly per thread per
// channel per stage.  Returns the buffer holding the result (always `b` for 9 stages).
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
        __syncthreads();
        float2 (*t)[N2] = src; src = dst; dst = t;
    }
}

} // namespace

__global__ void __launch_bounds__(256) convolve_chunk_kernel(
    const float* __restrict__ in, float* __restrict__ out, int stride, int offset, int n, int fill, int cur, int S, int nv,
    const float2* __restrict__ H, float2* __restrict__ fdl, float2* __restrict__ ypre,
    float* __restrict__ overlap, float* __restrict__ inbuf, const float2* __restrict__ twg) {
    extern __shared__ __align__(128) unsigned char smemRaw[];
    float2 (*stX)[CH][N2] = reinterpret_cast<float2 (*)[CH][N2]>(smemRaw);                                   // [STAGES][CH][512]
    float2 (*stH)[N2] = reinterpret_cast<float2 (*)[N2]>(smemRaw + (size_t) STAGES * CH * ROW_BYTES);        // [STAGES][512]
    float2 (*A)[N2] = reinterpret_cast<float2 (*)[N2]>(smemRaw + (size_t) STAGES * (CH + 1) * ROW_BYTES);    // [CH][512]
    float2 (*B)[N2] = A + CH;                                                                                 // [CH][512]
    float2* tw = reinterpret_cast<float2*>(B + CH);                                                           // [512]
    uint64_t* full = reinterpret_cast<uint64_t*>(tw + N2);                                                    // [STAGES]

    const int tid = threadIdx.x;
    const int ch0 = blockIdx.x * CH;

    tw[tid] = twg[tid];
    tw[tid + 256] = twg[tid + 256];

    // Producer side of the delay-line pipeline: one thread arms the stage's mbarrier with the byte count and issues
    // CH + 1 bulk copies (the CH channels' spectra of partition i and the IR spectrum H_i, 4 KB each).
    auto issueStage = [&](int i) {
        const int stg = (i - 1) % STAGES;
        int slotIdx = cur + i;
        if (slotIdx >= S) slotIdx -= S;
        mbar_expect_tx(&full[stg], (CH + 1) * ROW_BYTES);
#pragma unroll
        for (int c = 0; c < CH; ++c)
            tma_load_1d(&stX[stg][c][0], fdl + ((size_t) min(ch0 + c, nv - 1) * S + slotIdx) * NB, ROW_BYTES, &full[stg]);
        tma_load_1d(&stH[stg][0], H + (size_t) i * NB, ROW_BYTES, &full[stg]);
    };
    if (fill == 0 && tid == 0) {
        for (int sidx = 0; sidx < STAGES; ++sidx) mbar_init(&full[sidx], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
        for (int i = 1; i < S && i <= STAGES; ++i) issueStage(i);   // in flight while the forward FFT runs
    }

    // 1. append the chunk to the partition input buffer and load it as 512 complex points (even, odd), zero-padded
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int ch = ch0 + c;
        if (ch < nv) {
            float* ib = inbuf + (size_t) ch * CONV_BLOCK;
            for (int i = tid; i < n; i += 256) ib[fill + i] = in[(size_t) ch * stride + offset + i];
        }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int ch = ch0 + c;
        float2 z = make_float2(0.0f, 0.0f);
        if (ch < nv) z = reinterpret_cast<const float2*>(inbuf + (size_t) ch * CONV_BLOCK)[tid];
        A[c][tid] = z;
        A[c][tid + 256] = make_float2(0.0f, 0.0f);
    }
    __syncthreads();

    // 2. forward transform: Z = FFT512(z) lands in B; split into the bins of the real FFT, bin 512 packed into bin 0 (written to A)
    fft512<false>(A, B, tw, tid);
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
    __syncthreads();

    // 3./4. frequency-domain delay line: Y[b] = sum_i H_i[b] * X_{cur+i}[b]; the older partitions only once per block.
    // Bin 0 is the packed pair of real bins and multiplies component-wise.  Each thread owns bins tid and tid+256 of
    // CH channels; the loads are unconditional (padding channels alias the last real one) and unrolled so that
    // dozens of independent 8-byte loads are in flight per thread — the loop is pure HBM streaming.
    {
        float2 acc[2][CH];
        if (fill == 0) {
#pragma unroll
            for (int c = 0; c < CH; ++c) { acc[0][c] = make_float2(0.0f, 0.0f); acc[1][c] = make_float2(0.0f, 0.0f); }
            for (int i = 1; i < S; ++i) {
                const int stg = (i - 1) % STAGES;
                mbar_wait(&full[stg], ((i - 1) / STAGES) & 1);           // bytes of partition i have landed
                const float2 ha = stH[stg][tid], hb = stH[stg][tid + 256];
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const float2 xa = stX[stg][c][tid];
                    const float2 xb = stX[stg][c][tid + 256];
                    acc[0][c] = (tid == 0) ? make_float2(acc[0][c].x + ha.x * xa.x, acc[0][c].y + ha.y * xa.y) : cadd(acc[0][c], cmul(ha, xa));
                    acc[1][c] = cadd(acc[1][c], cmul(hb, xb));
                }
                __syncthreads();                                           // everyone is done with this stage ...
                if (tid == 0 && i + STAGES < S) { fence_proxy_async(); issueStage(i + STAGES); }   // ... refill it
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
        const float2 h0a = __ldg(H + tid), h0b = __ldg(H + tid + 256);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const float2 xa = A[c][tid], xb = A[c][tid + 256];
            B[c][tid] = (tid == 0) ? make_float2(acc[0][c].x + xa.x * h0a.x, acc[0][c].y + xa.y * h0a.y) : cadd(acc[0][c], cmul(xa, h0a));
            B[c][tid + 256] = cadd(acc[1][c], cmul(xb, h0b));
        }
    }
    __syncthreads();

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
    __syncthreads();
    fft512<true>(A, B, tw, tid);   // result in B: z[j] * 512

    // 5. overlap-add output for the samples of this chunk; save the second half when the partition is complete
    const float scale = 1.0f / 512.0f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int ch = ch0 + c;
        if (ch >= nv) continue;
        const float* y = reinterpret_cast<const float*>(&B[c][0]);      // y[2j] = re z[j], y[2j+1] = im z[j]
        float* ov = overlap + (size_t) ch * CONV_BLOCK;
        for (int i = tid; i < n; i += 256) out[(size_t) ch * stride + offset + i] = y[fill + i] * scale + ov[fill + i];
    }
    if (fill + n == CONV_BLOCK) {
        __syncthreads();
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int ch = ch0 + c;
            if (ch >= nv) continue;
            const float* y = reinterpret_cast<const float*>(&B[c][0]);
            float* ov = overlap + (size_t) ch * CONV_BLOCK;
            float* ib = inbuf + (size_t) ch * CONV_BLOCK;
            for (int i = tid; i < CONV_BLOCK; i += 256) { ov[i] = y[CONV_BLOCK + i] * scale; ib[i] = 0.0f; }
        }
    }
}

cudaError_t convolver_process_chunk(ConvolverState& st, const float* in, float* out, int stride, int offset, int n, cudaStream_t stream) {
    if (st.planOnly) return cudaErrorNotSupported;
    if (st.partitions == 0) {   // empty (fully trimmed) IR: silence (FFTConvolver.cpp:149-153)
        return cudaMemset2DAsync(out + offset, sizeof(float) * stride, 0, sizeof(float) * n, st.nv, stream);
    }
    const int grid = (st.nv + CONV_CH_PER_CTA - 1) / CONV_CH_PER_CTA;
    const size_t smem = (size_t) STAGES * (CH + 1) * ROW_BYTES + (size_t) 2 * CH * ROW_BYTES + ROW_BYTES + STAGES * sizeof(uint64_t);
    static bool attrSet = false;
    if (!attrSet) {
        cudaError_t e = cudaFuncSetAttribute(convolve_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e != cudaSuccess) return e;
        attrSet = true;
    }
    convolve_chunk_kernel<<<grid, 256, smem, stream>>>(in, out, stride, offset, n, st.fill, st.cur, st.partitions, st.nv,
                                                    st.dH, st.dFdl, st.dYpre, st.dOverlap, st.dInBuf, st.dTw);
    st.fill += n;
    if (st.fill == CONV_BLOCK) {
        st.fill = 0;
        st.cur = (st.cur > 0) ? st.cur - 1 : st.partitions - 1;   // FFTConvolver.cpp:200
    }
    return cudaGetLastError();
}

} // namespace eb
