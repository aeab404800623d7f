// render_kernel.cu — K1: the fused render-sequence kernel for sm_100a.
//
// Replaces, for every voice at once, the reference's per-block walk
//   GraphRenderSequence::process -> RootRenderSequence::process -> node->process(BlockContext)
//   (runtime/elem/GraphRenderSequence.h:268-309, :212-232)
// One warp owns one tile of up to 32 voices (lane == voice) for the whole block.  The block is cut into
// sample tiles of TILE samples; for each tile the warp interprets the compiled render program (program.h)
// op by op, warp-uniformly.  Each op handler runs the node's per-sample recurrence serially inside the lane
// for TILE samples (state in registers, carried across tiles in the warp's shared-memory state area and
// across blocks in HBM rows), reading its inputs from / writing its output to shared-memory slots
// [TILE][32] — the 2 KB-per-node block buffers of the reference never exist.  Independent samples of
// stateless ops (sin, tanh, mul, ...) are unrolled for ILP; only true recurrences are serial.
//
// Numerics: compiled with -fmad=false so a*b+c is two roundings exactly like the reference built with
// -ffp-contract=off; svf/svfshelf/mm1p/prewarp coefficient math is evaluated in double like the reference
// (filters/SVF.h:72-80); no -use_fast_math, no flush-to-zero.
//
// HBM layout (per voice group): rows[row][Vpad] f32 (params, scalar state; a double state is two rows viewed
// as double[Vpad]); delay rings / tap buffers [tile][pos][L] so equal write indices coalesce into one line.

#include <cuda_runtime.h>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include "program.h"
#include "kernels.h"

namespace eb {

namespace {

constexpr float kEps = FLT_EPSILON;

struct Opnd {
    const float* p;   // shared-memory slot column of this lane (valid only when slot)
    float k;          // broadcast value when not a slot
    bool slot;
};

template <int T>
struct Ctx {
    const LaunchParams* P;
    float* slots;     // [nSlots][T][L]
    float* outacc;    // [nOut][T][L]
    float* sst;       // [nStateRows][L]
    int lane;         // == column inside the tile; only lanes < tileWidth with a real voice stay alive
    int ls;           // lane stride of the shared-memory arrays (== tileWidth)
    int voice;        // voice index inside the group
    int tile;
    unsigned amask;   // mask of the live lanes of this warp
    int s0;           // first sample of the current tile
    int cnt;          // samples in the current tile (<= T)
};

#define LDO(o, t) ((o).slot ? (o).p[(t) * LS] : (o).k)

template <int T>
__device__ __forceinline__ Opnd decode(const Ctx<T>& c, uint32_t w) {
    Opnd o;
    const uint32_t kind = w >> 30, idx = w & 0x3FFFFFFFu;
    o.slot = (kind == K_SLOT);
    o.p = c.slots + (o.slot ? idx : 0u) * (T * c.ls) + c.lane;
    o.k = 0.0f;
    if (kind == K_PARAM) o.k = __ldg(c.P->rows + (size_t) idx * c.P->Vpad + c.voice);
    return o;
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }
__device__ __forceinline__ double clampd(double v, double lo, double hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }
__device__ __forceinline__ float stdmin(float a, float b) { return (b < a) ? b : a; }   // std::min
__device__ __forceinline__ float stdmax(float a, float b) { return (a < b) ? b : a; }   // std::max

// helpers/Change.h:12-32
__device__ __forceinline__ float change_tick(float& lastIn, float xn) {
    const float dt = xn - lastIn;
    lastIn = xn;
    return (dt > 0.0f) ? 1.0f : ((dt < 0.0f) ? -1.0f : 0.0f);
}

#define FOR_TILE(t) _Pragma("unroll") for (int t = 0; t < T; ++t) if (t < c.cnt)
#define FOR_TILE4(t) _Pragma("unroll 4") for (int t = 0; t < T; ++t) if (t < c.cnt)

template <int T, typename F>
__device__ __forceinline__ void map1(const Ctx<T>& c, const Opnd& a, float* out, F f) {
    const int LS = c.ls;
    FOR_TILE4(t) out[t * LS] = f(LDO(a, t));
}

// ---- Math.h:9-28 ---------------------------------------------------------------------------------------
template <int T>
__device__ void op_unary(const Ctx<T>& c, uint32_t mode, const Opnd& a, float* out) {
    switch (mode) {
        case U_SIN:   map1<T>(c, a, out, [](float x) { return sinf(x); }); break;
        case U_COS:   map1<T>(c, a, out, [](float x) { return cosf(x); }); break;
        case U_TAN:   map1<T>(c, a, out, [](float x) { return tanf(x); }); break;
        case U_TANH:  map1<T>(c, a, out, [](float x) { return tanhf(x); }); break;
        case U_ASINH: map1<T>(c, a, out, [](float x) { return asinhf(x); }); break;
        case U_LN:    map1<T>(c, a, out, [](float x) { return logf(x); }); break;
        case U_LOG10: map1<T>(c, a, out, [](float x) { return log10f(x); }); break;
        case U_LOG2:  map1<T>(c, a, out, [](float x) { return log2f(x); }); break;
        case U_CEIL:  map1<T>(c, a, out, [](float x) { return ceilf(x); }); break;
        case U_FLOOR: map1<T>(c, a, out, [](float x) { return floorf(x); }); break;
        case U_ROUND: map1<T>(c, a, out, [](float x) { return roundf(x); }); break;
        case U_SQRT:  map1<T>(c, a, out, [](float x) { return sqrtf(x); }); break;
        case U_EXP:   map1<T>(c, a, out, [](float x) { return expf(x); }); break;
        default:      map1<T>(c, a, out, [](float x) { return fabsf(x); }); break;
    }
}

// ---- Math.h:30-57,142-188 ------------------------------------------------------------------------------
__device__ __forceinline__ float binary_fn(uint32_t mode, float x, float y) {
    switch (mode) {
        case B_LE:  return (x < y) ? 1.0f : 0.0f;
        case B_LEQ: return (x <= y) ? 1.0f : 0.0f;
        case B_GE:  return (x > y) ? 1.0f : 0.0f;
        case B_GEQ: return (x >= y) ? 1.0f : 0.0f;
        case B_POW: return (x < 0.0f && y != floorf(y)) ? 0.0f : powf(x, y);
        case B_EQ:  return (fabsf(x - y) <= kEps) ? 1.0f : 0.0f;
        case B_AND: return (fabsf(1.0f - x) <= kEps && fabsf(1.0f - y) <= kEps) ? 1.0f : 0.0f;
        default:    return (fabsf(1.0f - x) <= kEps || fabsf(1.0f - y) <= kEps) ? 1.0f : 0.0f;
    }
}

template <int T>
__device__ void op_binary(const Ctx<T>& c, uint32_t mode, const Opnd& a, const Opnd& b, float* out) {
    const int LS = c.ls;
    FOR_TILE4(t) out[t * LS] = binary_fn(mode, LDO(a, t), LDO(b, t));
}

// ---- Math.h:59-89,128-177: left fold over the children in order ---------------------------------------------
__device__ __forceinline__ float reduce_fn(uint32_t mode, float x, float y) {
    switch (mode) {
        case R_ADD: return x + y;
        case R_SUB: return x - y;
        case R_MUL: return x * y;
        case R_DIV: return (y == 0.0f) ? 0.0f : x / y;
        case R_MOD: return fmodf(x, y);
        case R_MIN: return stdmin(x, y);
        default:    return stdmax(x, y);
    }
}

template <int T>
__device__ void op_reduce(const Ctx<T>& c, uint32_t mode, const uint32_t* opnds, int n, float* out) {
    const int LS = c.ls;
    float acc[T];
    {
        const Opnd a = decode<T>(c, __ldg(opnds));
        FOR_TILE(t) acc[t] = LDO(a, t);
    }
    for (int j = 1; j < n; ++j) {
        const Opnd b = decode<T>(c, __ldg(opnds + j));
        switch (mode) {   // hoisted so the inner loop is branch-free
            case R_ADD: FOR_TILE(t) acc[t] = acc[t] + LDO(b, t); break;
            case R_SUB: FOR_TILE(t) acc[t] = acc[t] - LDO(b, t); break;
            case R_MUL: FOR_TILE(t) acc[t] = acc[t] * LDO(b, t); break;
            default:    FOR_TILE(t) acc[t] = reduce_fn(mode, acc[t], LDO(b, t)); break;
        }
    }
    FOR_TILE(t) out[t * LS] = acc[t];
}

} // namespace

// =========================================================================================================
template <int T>
__global__ void __launch_bounds__(256) render_block_kernel(const __grid_constant__ LaunchParams P) {
    extern __shared__ __align__(16) float smem[];

    const int warpsPerCta = blockDim.x >> 5;
    const int warpInCta = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int L = P.tileWidth;
    const int LS = L;
    const int nTiles = (P.nv + L - 1) / L;
    const int tile = blockIdx.x * warpsPerCta + warpInCta;
    if (tile >= nTiles) return;   // whole warp leaves together

    // Only lanes that own a real voice stay alive (tile width L <= 32; the last tile may be ragged).  The host
    // shrinks L when there are few voices so that more warps — and more SMs — work on them.
    const int count = min(L, P.nv - tile * L);
    if (lane >= count) return;

    Ctx<T> c;
    c.P = &P;
    c.lane = lane;
    c.ls = LS;
    c.tile = tile;
    c.voice = tile * L + lane;
    c.amask = (count >= 32) ? 0xFFFFFFFFu : ((1u << count) - 1u);
    const int perWarp = ((P.nSlots * T + P.nOut * T + P.nStateRows) * LS + 3) & ~3;
    c.slots = smem + (size_t) warpInCta * perWarp;
    c.outacc = c.slots + P.nSlots * T * LS;
    c.sst = c.outacc + P.nOut * T * LS;

    // ---- state rows HBM -> shared memory (once per block) ----
    {
        int srow = 0;
        for (int e = 0; e < P.nStateEntries; ++e) {
            const uint32_t m = __ldg(P.stateMap + e);
            if (m == STATE_PAD) { srow += 1; continue; }
            const size_t row = m & ~STATE_DOUBLE_FLAG;
            if (m & STATE_DOUBLE_FLAG) {
                const double* g = reinterpret_cast<const double*>(P.rows + row * P.Vpad);
                reinterpret_cast<double*>(c.sst + srow * LS)[lane] = g[c.voice];
                srow += 2;
            } else {
                c.sst[srow * LS + lane] = P.rows[row * P.Vpad + c.voice];
                srow += 1;
            }
        }
    }

    const int numSamples = P.numSamples;
    for (int s0 = 0; s0 < numSamples; s0 += T) {
        c.s0 = s0;
        c.cnt = min(T, numSamples - s0);

        for (int i = 0; i < P.nOut * T; ++i) c.outacc[i * LS + lane] = 0.0f;

        const uint32_t* pc = P.code;
        for (;;) {
            const uint32_t w0 = __ldg(pc);
            const uint32_t opcode = w0 & 0xFF, nopnd = (w0 >> 8) & 0xFF, mode = w0 >> 24;
            if (opcode == OP_END) break;
            float* out = c.slots + ((w0 >> 16) & 0xFF) * (T * LS) + lane;
            const uint32_t sidx = __ldg(pc + 1);
            const uint32_t aux0 = __ldg(pc + 2), aux1 = __ldg(pc + 3);
            const uint32_t* opnds = pc + OP_HEADER_WORDS;
            const uint64_t ptrbits = (uint64_t) __ldg(pc + 4) | ((uint64_t) __ldg(pc + 5) << 32);
            pc += OP_HEADER_WORDS + nopnd;

            switch (opcode) {
            case OP_SEG:
                if (!((P.runMask >> aux0) & 1u)) pc += aux1;
                break;

            case OP_FILL0:
                FOR_TILE(t) out[t * LS] = 0.0f;
                break;

            case OP_COPY: {
                const Opnd a = decode<T>(c, __ldg(opnds));
                FOR_TILE(t) out[t * LS] = LDO(a, t);
            } break;

            case OP_LOADIN: {
                if (P.inVoice) {
                    const float* g = P.inVoice + ((size_t) (P.voice0 + c.voice) * P.nIn + aux0) * P.inStride + s0;
                    FOR_TILE(t) out[t * LS] = g[t];
                } else {
                    const float* g = P.inShared + (size_t) aux0 * P.inStride + s0;
                    FOR_TILE(t) out[t * LS] = __ldg(g + t);
                }
            } break;

            case OP_UNARY: {
                const Opnd a = decode<T>(c, __ldg(opnds));
                op_unary<T>(c, mode, a, out);
            } break;

            case OP_BINARY: {
                const Opnd a = decode<T>(c, __ldg(opnds));
                const Opnd b = decode<T>(c, __ldg(opnds + 1));
                op_binary<T>(c, mode, a, b, out);
            } break;

            case OP_REDUCE:
                op_reduce<T>(c, mode, opnds, (int) nopnd, out);
                break;

            case OP_PHASOR: {   // Core.h:89-97: step = f * (1/sr) in float; phase = next - floor(next)
                const Opnd f = decode<T>(c, __ldg(opnds));
                const float rsr = __uint_as_float(aux0);
                float phase = c.sst[sidx * LS + lane];
                FOR_TILE(t) {
                    const float step = LDO(f, t) * rsr;
                    out[t * LS] = phase;
                    const float next = phase + step;
                    phase = next - floorf(next);
                }
                c.sst[sidx * LS + lane] = phase;
            } break;

            case OP_SPHASOR: {  // Core.h:113-121; state: phase, change.lastIn
                const Opnd f = decode<T>(c, __ldg(opnds));
                const Opnd r = decode<T>(c, __ldg(opnds + 1));
                const float rsr = __uint_as_float(aux0);
                float phase = c.sst[sidx * LS + lane];
                float last = c.sst[(sidx + 1) * LS + lane];
                FOR_TILE(t) {
                    const float xn = LDO(f, t);
                    if (change_tick(last, LDO(r, t)) > 0.5f) phase = 0.0f;
                    const float step = xn * rsr;
                    out[t * LS] = phase;
                    const float next = phase + step;
                    phase = next - floorf(next);
                }
                c.sst[sidx * LS + lane] = phase;
                c.sst[(sidx + 1) * LS + lane] = last;
            } break;

            case OP_COUNTER: {  // Core.h:198-211
                const Opnd g = decode<T>(c, __ldg(opnds));
                float count = c.sst[sidx * LS + lane];
                FOR_TILE(t) {
                    const float in = LDO(g, t);
                    if ((1.0f - in) <= kEps) { out[t * LS] = count; count = count + 1.0f; }
                    else { count = 0.0f; out[t * LS] = 0.0f; }
                }
                c.sst[sidx * LS + lane] = count;
            } break;

            case OP_ACCUM: {    // Core.h:233-243; state: runningTotal, change.lastIn
                const Opnd x = decode<T>(c, __ldg(opnds));
                const Opnd r = decode<T>(c, __ldg(opnds + 1));
                float total = c.sst[sidx * LS + lane];
                float last = c.sst[(sidx + 1) * LS + lane];
                FOR_TILE(t) {
                    if (change_tick(last, LDO(r, t)) > 0.5f) total = 0.0f;
                    total += LDO(x, t);
                    out[t * LS] = total;
                }
                c.sst[sidx * LS + lane] = total;
                c.sst[(sidx + 1) * LS + lane] = last;
            } break;

            case OP_LATCH: {    // Core.h:265-281; state: z, hold
                const Opnd l = decode<T>(c, __ldg(opnds));
                const Opnd x = decode<T>(c, __ldg(opnds + 1));
                float z = c.sst[sidx * LS + lane];
                float hold = c.sst[(sidx + 1) * LS + lane];
                FOR_TILE(t) {
                    const float lv = LDO(l, t);
                    if (fabsf(z) <= kEps && lv > kEps) hold = LDO(x, t);
                    z = lv;
                    out[t * LS] = hold;
                }
                c.sst[sidx * LS + lane] = z;
                c.sst[(sidx + 1) * LS + lane] = hold;
            } break;

            case OP_MAXHOLD: {  // Core.h:315-332; state: max, samplesAtCurrentMax(u32), change.lastIn; aux0 = holdTimeSamples
                const Opnd x = decode<T>(c, __ldg(opnds));
                const Opnd r = decode<T>(c, __ldg(opnds + 1));
                float mx = c.sst[sidx * LS + lane];
                uint32_t held = __float_as_uint(c.sst[(sidx + 1) * LS + lane]);
                float last = c.sst[(sidx + 2) * LS + lane];
                const uint32_t hts = aux0;
                FOR_TILE(t) {
                    const float in = LDO(x, t);
                    bool reset = change_tick(last, LDO(r, t)) > 0.5f;
                    if (!reset) reset = (++held >= hts);   // short-circuit || of the reference
                    if (reset) { mx = in; held = 0; }
                    else if (in > mx) { held = 0; mx = in; }
                    out[t * LS] = mx;
                }
                c.sst[sidx * LS + lane] = mx;
                c.sst[(sidx + 1) * LS + lane] = __uint_as_float(held);
                c.sst[(sidx + 2) * LS + lane] = last;
            } break;

            case OP_RAND: {     // Noise.h:25-38
                uint32_t seed = __float_as_uint(c.sst[sidx * LS + lane]);
                FOR_TILE(t) {
                    seed = 214013u * seed + 2531011u;
                    const int r = (int) ((seed >> 16) & 0x7FFFu);
                    out[t * LS] = (float) r / 32767.0f;
                }
                c.sst[sidx * LS + lane] = __uint_as_float(seed);
            } break;

            case OP_POLE: {     // Filters.h:27-33
                const Opnd pp = decode<T>(c, __ldg(opnds));
                const Opnd x = decode<T>(c, __ldg(opnds + 1));
                float z = c.sst[sidx * LS + lane];
                FOR_TILE(t) {
                    z = LDO(x, t) + LDO(pp, t) * z;
                    out[t * LS] = z;
                }
                c.sst[sidx * LS + lane] = z;
            } break;

            case OP_ENV: {      // Filters.h:61-73
                const Opnd ap = decode<T>(c, __ldg(opnds));
                const Opnd rp = decode<T>(c, __ldg(opnds + 1));
                const Opnd x = decode<T>(c, __ldg(opnds + 2));
                float z = c.sst[sidx * LS + lane];
                FOR_TILE(t) {
                    const float vn = fabsf(LDO(x, t));
                    const float pcoef = (vn > z) ? LDO(ap, t) : LDO(rp, t);
                    z = pcoef * (z - vn) + vn;
                    out[t * LS] = z;
                }
                c.sst[sidx * LS + lane] = z;
            } break;

            case OP_BIQUAD: {   // Filters.h:102-114 (TDF-II, audio-rate coefficients)
                const Opnd b0 = decode<T>(c, __ldg(opnds));
                const Opnd b1 = decode<T>(c, __ldg(opnds + 1));
                const Opnd b2 = decode<T>(c, __ldg(opnds + 2));
                const Opnd a1 = decode<T>(c, __ldg(opnds + 3));
                const Opnd a2 = decode<T>(c, __ldg(opnds + 4));
                const Opnd x = decode<T>(c, __ldg(opnds + 5));
                float z1 = c.sst[sidx * LS + lane];
                float z2 = c.sst[(sidx + 1) * LS + lane];
                FOR_TILE(t) {
                    const float xn = LDO(x, t);
                    const float y = LDO(b0, t) * xn + z1;
                    z1 = LDO(b1, t) * xn - LDO(a1, t) * y + z2;
                    z2 = LDO(b2, t) * xn - LDO(a2, t) * y;
                    out[t * LS] = y;
                }
                c.sst[sidx * LS + lane] = z1;
                c.sst[(sidx + 1) * LS + lane] = z2;
            } break;

            case OP_PREWARP: {  // filters/MultiMode1p.h:23-33; (aux0,aux1) = bits of T = 1.0/sr (double)
                const Opnd fc = decode<T>(c, __ldg(opnds));
                const double Ts = __longlong_as_double((long long) ((uint64_t) aux0 | ((uint64_t) aux1 << 32)));
                FOR_TILE4(t) {
                    const double twoPi = 2.0 * 3.141592653589793238;
                    const double wd = twoPi * (double) LDO(fc, t);
                    out[t * LS] = (float) tan(wd * Ts / 2.0);
                }
            } break;

            case OP_MM1P: {     // filters/MultiMode1p.h:78-103; state: double z; mode 0 low / 2 high / 4 all
                const Opnd gi = decode<T>(c, __ldg(opnds));
                const Opnd x = decode<T>(c, __ldg(opnds + 1));
                double* zs = reinterpret_cast<double*>(c.sst + sidx * LS) + lane;
                double z = *zs;
                FOR_TILE(t) {
                    const double g = clampd((double) LDO(gi, t), 0.0, 0.9999);
                    const float xn = LDO(x, t);
                    const double G = g / (1.0 + g);
                    const double v = ((double) xn - z) * G;
                    const double lp = v + z;
                    z = lp + v;
                    float y;
                    if (mode == 0) y = (float) lp;
                    else if (mode == 2) y = xn - (float) lp;
                    else y = (float) (lp + lp - (double) xn);
                    out[t * LS] = y;
                }
                *zs = z;
            } break;

            case OP_SVF: {      // filters/SVF.h:48-104; (aux0,aux1) = bits of sr (double); state: double ic1eq, ic2eq
                const Opnd fc = decode<T>(c, __ldg(opnds));
                const Opnd q = decode<T>(c, __ldg(opnds + 1));
                const Opnd x = decode<T>(c, __ldg(opnds + 2));
                const double sr = __longlong_as_double((long long) ((uint64_t) aux0 | ((uint64_t) aux1 << 32)));
                const double fmax = sr / 2.0001;
                double* s1 = reinterpret_cast<double*>(c.sst + sidx * LS) + lane;
                double* s2 = reinterpret_cast<double*>(c.sst + (sidx + 2) * LS) + lane;
                double ic1 = *s1, ic2 = *s2;
                // coefficient math is independent per sample: do it first (ILP), then the serial tick
                double a1[T], a2[T], a3[T], kk[T];
                FOR_TILE(t) {
                    const double g = tan(3.14159265359 * clampd((double) LDO(fc, t), 20.0, fmax) / sr);
                    const double k = 1.0 / clampd((double) LDO(q, t), 0.25, 20.0);
                    a1[t] = 1.0 / (1.0 + g * (g + k));
                    a2[t] = g * a1[t];
                    a3[t] = g * a2[t];
                    kk[t] = k;
                }
                FOR_TILE(t) {
                    const float v0 = LDO(x, t);
                    const double v3 = (double) v0 - ic2;
                    const double v1 = ic1 * a1[t] + v3 * a2[t];
                    const double v2 = ic2 + ic1 * a2[t] + v3 * a3[t];
                    ic1 = v1 * 2.0 - ic1;
                    ic2 = v2 * 2.0 - ic2;
                    float y;
                    switch (mode) {
                        case 0: y = (float) v2; break;
                        case 1: y = (float) v1; break;
                        case 2: y = (float) ((double) v0 - kk[t] * v1 - v2); break;
                        case 3: y = (float) ((double) v0 - kk[t] * v1); break;
                        default: y = (float) ((double) v0 - 2.0 * kk[t] * v1); break;
                    }
                    out[t * LS] = y;
                }
                *s1 = ic1; *s2 = ic2;
            } break;

            case OP_SVFSHELF: { // filters/SVFShelf.h:44-106; mode 0 lowshelf / 1 highshelf / 2 bell
                const Opnd fc = decode<T>(c, __ldg(opnds));
                const Opnd q = decode<T>(c, __ldg(opnds + 1));
                const Opnd gdb = decode<T>(c, __ldg(opnds + 2));
                const Opnd x = decode<T>(c, __ldg(opnds + 3));
                const double sr = __longlong_as_double((long long) ((uint64_t) aux0 | ((uint64_t) aux1 << 32)));
                const double fmax = sr / 2.0001;
                double* s1 = reinterpret_cast<double*>(c.sst + sidx * LS) + lane;
                double* s2 = reinterpret_cast<double*>(c.sst + (sidx + 2) * LS) + lane;
                double ic1 = *s1, ic2 = *s2;
                _Pragma("unroll 2") for (int t = 0; t < T; ++t) if (t < c.cnt) {
                    const double A = pow(10.0, (double) LDO(gdb, t) / 40.0);
                    double g = tan(3.14159265359 * clampd((double) LDO(fc, t), 20.0, fmax) / sr);
                    double k = 1.0 / clampd((double) LDO(q, t), 0.25, 20.0);
                    if (mode == 0) g /= A;
                    if (mode == 1) g *= A;
                    if (mode == 2) k /= A;
                    const double a1 = 1.0 / (1.0 + g * (g + k));
                    const double a2 = g * a1;
                    const double a3 = g * a2;
                    const float v0 = LDO(x, t);
                    const double v3 = (double) v0 - ic2;
                    const double v1 = ic1 * a1 + v3 * a2;
                    const double v2 = ic2 + ic1 * a2 + v3 * a3;
                    ic1 = v1 * 2.0 - ic1;
                    ic2 = v2 * 2.0 - ic2;
                    float y;
                    if (mode == 2) y = (float) ((double) v0 + k * (A * A - 1.0) * v1);
                    else if (mode == 0) y = (float) ((double) v0 + k * (A - 1.0) * v1 + (A * A - 1.0) * v2);
                    else y = (float) (A * A * (double) v0 + k * (1.0 - A) * A * v1 + (1.0 - A * A) * v2);
                    out[t * LS] = y;
                }
                *s1 = ic1; *s2 = ic2;
            } break;

            case OP_Z: {        // Delays.h:29-34
                const Opnd x = decode<T>(c, __ldg(opnds));
                float z = c.sst[sidx * LS + lane];
                FOR_TILE(t) { out[t * LS] = z; z = LDO(x, t); }
                c.sst[sidx * LS + lane] = z;
            } break;

            case OP_DELAY: {    // Delays.h:108-159; aux0 = size; ring [tile][pos][L]; state: writeIndex
                const Opnd len = decode<T>(c, __ldg(opnds));
                const Opnd fb = decode<T>(c, __ldg(opnds + 1));
                const Opnd x = decode<T>(c, __ldg(opnds + 2));
                const int size = (int) aux0;
                if (size == 0) { FOR_TILE(t) out[t * LS] = LDO(len, t); break; }   // Delays.h:105-106 copies inputData[0]
                float* ring = reinterpret_cast<float*>(ptrbits) + (size_t) tile * size * L + lane;
                int w = __float_as_int(c.sst[sidx * LS + lane]);
                const float fsize = (float) size;
                FOR_TILE(t) {
                    const float offset = clampf(LDO(len, t), 0.0f, fsize);
                    float y, in;
                    if (offset <= kEps) {
                        in = LDO(x, t);
                        y = in;
                    } else {
                        const float readFrac = (float) (size + w) - offset;
                        const int readLeft = (int) readFrac;
                        const int readRight = readLeft + 1;
                        const float frac = readFrac - floorf(readFrac);
                        const float left = ring[(size_t) (readLeft % size) * L];
                        const float right = ring[(size_t) (readRight % size) * L];
                        y = left + frac * (right - left);
                        const float fbv = clampf(LDO(fb, t), -1.0f, 1.0f);
                        in = LDO(x, t) + fbv * y;
                    }
                    ring[(size_t) w * L] = in;
                    out[t * LS] = y;
                    if (++w >= size) w -= size;
                }
                c.sst[sidx * LS + lane] = __int_as_float(w);
            } break;

            case OP_SDELAY: {   // Delays.h:246-260; aux0 = ring size (pow2), aux1 = length; state: writeIndex
                const Opnd x = decode<T>(c, __ldg(opnds));
                const int size = (int) aux0, mask = size - 1, len = (int) aux1;
                float* ring = reinterpret_cast<float*>(ptrbits) + (size_t) tile * size * L + lane;
                int w = __float_as_int(c.sst[sidx * LS + lane]);
                // block-write-then-read == per-sample write-then-read because size >= len + blockSize
                FOR_TILE(t) {
                    const float in = LDO(x, t);
                    ring[(size_t) w * L] = in;
                    out[t * LS] = (len == 0) ? in : ring[(size_t) ((size + w - len) & mask) * L];
                    w = (w + 1) & mask;
                }
                c.sst[sidx * LS + lane] = __int_as_float(w);
            } break;

            case OP_TABLE: {    // Table.h:59-71; aux0 = table length; ptr = device copy of resource channel 0
                const Opnd pos = decode<T>(c, __ldg(opnds));
                const int size = (int) aux0;
                const float* tab = reinterpret_cast<const float*>(ptrbits);
                FOR_TILE(t) {
                    const float readPos = clampf(LDO(pos, t), 0.0f, 1.0f) * (float) (size - 1);
                    const int readLeft = (int) readPos;
                    const int readRight = readLeft + 1;
                    const float frac = readPos - floorf(readPos);
                    const float left = __ldg(tab + (readLeft % size));
                    const float right = __ldg(tab + (readRight % size));
                    out[t * LS] = left + frac * (right - left);
                }
            } break;

            case OP_BLEP: {     // Oscillators.h:23-89; state: phase, acc; aux0 = bits of float(sr)
                const Opnd f = decode<T>(c, __ldg(opnds));
                const float sr = __uint_as_float(aux0);
                float phase = c.sst[sidx * LS + lane];
                float acc = c.sst[(sidx + 1) * LS + lane];
                auto blep = [](float ph, float inc) -> float {
                    if (ph < inc) { const float p = ph / inc; return (2.0f - p) * p - 1.0f; }
                    if (ph > (1.0f - inc)) { const float p = (ph - 1.0f) / inc; return (p + 2.0f) * p + 1.0f; }
                    return 0.0f;
                };
                FOR_TILE(t) {
                    const float inc = LDO(f, t) / sr;
                    float y;
                    if (mode == 0) {
                        y = 2.0f * phase - 1.0f - blep(phase, inc);
                    } else {
                        const float naive = (phase < 0.5f) ? 1.0f : -1.0f;
                        const float halfPhase = fmodf(phase + 0.5f, 1.0f);
                        const float square = naive + blep(phase, inc) - blep(halfPhase, inc);
                        if (mode == 1) y = square;
                        else { acc += 4.0f * inc * square; y = acc; }
                    }
                    phase += inc;
                    if (phase >= 1.0f) phase -= 1.0f;
                    out[t * LS] = y;
                }
                c.sst[sidx * LS + lane] = phase;
                c.sst[(sidx + 1) * LS + lane] = acc;
            } break;

            case OP_TAPIN: {    // Feedback.h:42-52; ptr = shared tap buffer [tile][blockSize][L]
                const float* tap = reinterpret_cast<const float*>(ptrbits) + (size_t) tile * P.blockSize * L + lane;
                FOR_TILE(t) out[t * LS] = tap[(size_t) (s0 + t) * L];
            } break;

            case OP_TAPOUT: {   // Feedback.h:109-121; ptr = this node's private delayBuffer [tile][blockSize][L]
                const Opnd x = decode<T>(c, __ldg(opnds));
                float* buf = reinterpret_cast<float*>(ptrbits) + (size_t) tile * P.blockSize * L + lane;
                FOR_TILE(t) {
                    const float v = LDO(x, t);
                    buf[(size_t) (s0 + t) * L] = v;
                    out[t * LS] = v;
                }
            } break;

            case OP_STOREBUF: { // stage boundary: ptr = [voice][blockSize] staging buffer
                const Opnd x = decode<T>(c, __ldg(opnds));
                float* buf = reinterpret_cast<float*>(ptrbits) + (size_t) c.voice * P.blockSize + s0;
                FOR_TILE(t) buf[t] = LDO(x, t);
            } break;

            case OP_LOADBUF: {
                const float* buf = reinterpret_cast<const float*>(ptrbits) + (size_t) c.voice * P.blockSize + s0;
                FOR_TILE(t) out[t * LS] = buf[t];
            } break;

            case OP_ROOT: {     // Core.h:66-78 + GainFade.h:56-72; aux0 = root index
                const RootDyn rd = P.roots[aux0];
                if (nopnd < 1) { FOR_TILE(t) out[t * LS] = 0.0f; }
                else {
                    const Opnd x = decode<T>(c, __ldg(opnds));
                    if (rd.gain0 == rd.target) {
                        FOR_TILE(t) out[t * LS] = LDO(x, t) * rd.target;
                    } else {
                        FOR_TILE(t) {
                            const float g = clampf(rd.gain0 + rd.step * (float) (s0 + t), 0.0f, 1.0f);
                            out[t * LS] = LDO(x, t) * g;
                        }
                    }
                }
                if (rd.channel >= 0 && rd.channel < P.nOut) {   // GraphRenderSequence.h:227-231
                    float* acc = c.outacc + rd.channel * (T * LS) + lane;
                    FOR_TILE(t) acc[t * LS] += out[t * LS];
                }
            } break;

            default: break;
            }
        }

        // ---- tile epilogue: per-voice output and per-tile partial mix ----
        if (P.outVoice) {
            for (int ch = 0; ch < P.nOut; ++ch) {
                float* g = P.outVoice + ((size_t) (P.voice0 + c.voice) * P.nOut + ch) * P.outStride + s0;
                const float* a = c.outacc + ch * (T * LS) + lane;
                if (c.cnt == T && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
                    _Pragma("unroll") for (int t = 0; t < T; t += 4)
                        *reinterpret_cast<float4*>(g + t) = make_float4(a[t * LS], a[(t + 1) * LS], a[(t + 2) * LS], a[(t + 3) * LS]);
                } else {
                    FOR_TILE(t) g[t] = a[t * LS];
                }
            }
        }
        if (P.mixPartial) {
            // sum over the live lanes of the tile in a fixed order (deterministic), lane 0 publishes
            for (int ch = 0; ch < P.nOut; ++ch) {
                const float* a = c.outacc + ch * (T * LS) + lane;
                float* gp = P.mixPartial + ((size_t) (P.tileBase + tile) * P.nOut + ch) * P.blockSize + s0;
                _Pragma("unroll") for (int t = 0; t < T; ++t) {
                    float v = a[t * LS];
                    _Pragma("unroll") for (int d = 16; d > 0; d >>= 1) {
                        const float o = __shfl_down_sync(c.amask, v, d);
                        if (lane + d < count) v += o;
                    }
                    if (lane == 0 && t < c.cnt) gp[t] = v;
                }
            }
        }
        __syncwarp(c.amask);
    }

    // ---- state rows shared memory -> HBM ----
    {
        int srow = 0;
        for (int e = 0; e < P.nStateEntries; ++e) {
            const uint32_t m = __ldg(P.stateMap + e);
            if (m == STATE_PAD) { srow += 1; continue; }
            const size_t row = m & ~STATE_DOUBLE_FLAG;
            if (m & STATE_DOUBLE_FLAG) {
                double* g = reinterpret_cast<double*>(P.rows + row * P.Vpad);
                g[c.voice] = reinterpret_cast<const double*>(c.sst + srow * LS)[lane];
                srow += 2;
            } else {
                P.rows[row * P.Vpad + c.voice] = c.sst[srow * LS + lane];
                srow += 1;
            }
        }
    }

    // ---- tap promotion (GraphRenderSequence.h:200-210,306-308): ops after OP_END, until the second OP_END ----
    {
        const uint32_t* pc = P.code;
        for (;;) {   // skip the main program
            const uint32_t w0 = __ldg(pc);
            if ((w0 & 0xFF) == OP_END) { pc += 1; break; }
            pc += OP_HEADER_WORDS + ((w0 >> 8) & 0xFF);
        }
        for (;;) {
            const uint32_t w0 = __ldg(pc);
            if ((w0 & 0xFF) == OP_END) break;
            // promote record: [w0][root index][src lo][src hi][dst lo][dst hi]
            const uint32_t r = __ldg(pc + 1);
            const uint64_t sb = (uint64_t) __ldg(pc + 2) | ((uint64_t) __ldg(pc + 3) << 32);
            const uint64_t db = (uint64_t) __ldg(pc + 4) | ((uint64_t) __ldg(pc + 5) << 32);
            pc += OP_HEADER_WORDS;
            // only roots that are still the active target promote (RootRenderSequence::promoteTapBuffers)
            if (!((P.runMask >> (16 + r)) & 1u)) continue;
            const float* src = reinterpret_cast<const float*>(sb) + (size_t) tile * P.blockSize * L + lane;
            float* dst = reinterpret_cast<float*>(db) + (size_t) tile * P.blockSize * L + lane;
            for (int s = 0; s < numSamples; ++s) dst[(size_t) s * L] = src[(size_t) s * L];
        }
    }
}

// ---- deterministic reduction of the per-tile partial mixes: out[ch][s] = sum over tiles in fixed order ----
__global__ void __launch_bounds__(1024) mix_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                          int nTiles, int nOut, int blockSize, int numSamples) {
    __shared__ float red[32][33];
    const int sx = threadIdx.x & 31, gy = threadIdx.x >> 5;       // 32 samples x 32 tile-groups
    const int chunksPerCh = (blockSize + 31) / 32;
    const int ch = blockIdx.x / chunksPerCh;
    const int s = (blockIdx.x % chunksPerCh) * 32 + sx;
    float acc = 0.0f;
    if (s < numSamples)
        for (int t = gy; t < nTiles; t += 32) acc += partial[((size_t) t * nOut + ch) * blockSize + s];
    red[gy][sx] = acc;
    __syncthreads();
    if (gy == 0 && s < numSamples) {
        float v = 0.0f;
        for (int g = 0; g < 32; ++g) v += red[g][sx];
        out[(size_t) ch * blockSize + s] = v;
    }
}

// =========================================================================================================
// host-side launchers (called from graph_host.cpp)

size_t render_smem_bytes(int tileSamples, int nSlots, int nOut, int nStateRows, int warpsPerCta, int tileWidth) {
    const size_t perWarp = ((size_t) (nSlots * tileSamples + nOut * tileSamples + nStateRows) * tileWidth + 3) & ~(size_t) 3;
    return (size_t) warpsPerCta * perWarp * sizeof(float);
}

cudaError_t launch_render_block(const LaunchParams& P, int tileSamples, int warpsPerCta, cudaStream_t stream) {
    const int L = P.tileWidth;
    const int nTiles = (P.nv + L - 1) / L;
    if (nTiles <= 0) return cudaSuccess;
    const int grid = (nTiles + warpsPerCta - 1) / warpsPerCta;
    const size_t smem = render_smem_bytes(tileSamples, P.nSlots, P.nOut, P.nStateRows, warpsPerCta, P.tileWidth);
    cudaError_t e;
    if (tileSamples == 8) {
        e = cudaFuncSetAttribute(render_block_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e != cudaSuccess) return e;
        render_block_kernel<8><<<grid, warpsPerCta * 32, smem, stream>>>(P);
    } else {
        e = cudaFuncSetAttribute(render_block_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e != cudaSuccess) return e;
        render_block_kernel<4><<<grid, warpsPerCta * 32, smem, stream>>>(P);
    }
    return cudaGetLastError();
}

cudaError_t launch_mix_reduce(const float* partial, float* out, int nTiles, int nOut, int blockSize, int numSamples, cudaStream_t stream) {
    const int chunksPerCh = (blockSize + 31) / 32;
    mix_reduce_kernel<<<nOut * chunksPerCh, 1024, 0, stream>>>(partial, out, nTiles, nOut, blockSize, numSamples);
    return cudaGetLastError();
}

} // namespace eb
