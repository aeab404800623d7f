// render_kernel.cu — K1: the fused render-sequence kernel for sm_100a.
//
// Replaces, for every voice at once, the reference's per-block walk
//   GraphRenderSequence::process -> RootRenderSequence::process -> node->process(BlockContext)
//   (runtime/elem/GraphRenderSequence.h:268-309, :212-232)
//
// Work decomposition.  One warp owns one *voice tile* of L voices (L = 1..32, chosen by the host so that few
// voices still fill the machine) for the whole block.  The block is cut into sample tiles of T samples with
// E = L*T = 32*NITER elements; a node output for one tile is a shared-memory slot [T][L] (element e = t*L + v)
// — the reference's 2 KB-per-node block buffers never exist.  For each sample tile the warp interprets the
// compiled render program (program.h) op by op, warp-uniformly:
//   * stateless ops (math, compare, fades, table lookups, prewarp, svf coefficient math, delay lines whose read
//     head is outside the tile, ...) run with ALL 32 lanes over the E elements of the tile — lanes are voices
//     when L = 32 and consecutive samples of one voice when L = 1, the code is the same;
//   * true recurrences (phasor, svf tick, pole, biquad, ...) run serially over the T samples inside the lane that
//     owns the voice (lanes < L), state in registers, carried across tiles in the warp's shared-memory state
//     area and across blocks in HBM rows.
// So with few voices the time axis of everything that is not a recurrence is spread over the lanes, and with
// many voices every lane is a voice; results are identical either way because no floating-point operation is
// re-associated.
//
// Numerics: compiled with -fmad=false so a*b+c is two roundings exactly like the reference built with
// -ffp-contract=off; svf/svfshelf/mm1p/prewarp coefficient math is evaluated in double like the reference
// (filters/SVF.h:72-80); no -use_fast_math, no flush-to-zero.
//
// HBM layout (per voice group): rows[row][Vpad] f32 (params, scalar state; a double state is two rows viewed
// as double[Vpad]); delay rings / tap buffers [tile][pos][L] so the lanes of a warp touch one contiguous line.

#include <cuda_runtime.h>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include "program.h"
#include "kernels.h"

namespace eb {

namespace {

constexpr float kEps = FLT_EPSILON;
constexpr unsigned FULL = 0xFFFFFFFFu;

struct Opnd {
    const float* p;   // shared-memory slot base (valid only when slot)
    float k;          // this lane's voice parameter when not a slot
    bool slot;
};

struct Ctx {
    const LaunchParams* P;
    float* slots;     // [nSlots][E]
    float* outacc;    // [nOut][E]
    float* sst;       // [nStateRows][L]
    int lane;
    int vlane;        // lane & (L-1): the voice column this lane works for
    int L, logL, E, T;
    int voice;        // tile*L + vlane (may be a padding voice >= nv: it owns storage but is never output)
    int tile;
    bool valid;       // voice < nv
    bool owner;       // lane < L: this lane runs the recurrences of voice `voice`
    int s0, cnt;      // current sample tile: first sample, number of samples (<= T)
};

#define LDE(o, e) ((o).slot ? (o).p[(e)] : (o).k)                       /* element access (stateless ops) */
#define LDT(o, t) ((o).slot ? (o).p[(t) * c.L + c.lane] : (o).k)        /* sample access of the owner lane */

__device__ __forceinline__ Opnd decode(const Ctx& c, uint32_t w) {
    Opnd o;
    const uint32_t kind = w >> 30, idx = w & 0x3FFFFFFFu;
    o.slot = (kind == K_SLOT);
    o.p = c.slots + (o.slot ? idx : 0u) * c.E;
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

// Stateless element loop: all 32 lanes, NITER elements each.
#define FOR_ELEM(k, e, t)                                                                    \
    _Pragma("unroll") for (int k = 0; k < NITER; ++k) {                                      \
        const int e = c.lane + 32 * k;                                                       \
        const int t = e >> c.logL;                                                           \
        if (t < c.cnt) {
#define END_ELEM }}

// Serial sample loop of the lane that owns the voice.
#define FOR_OWNER(t) _Pragma("unroll 4") for (int t = 0; t < c.cnt; ++t)

template <int NITER, typename F>
__device__ __forceinline__ void map1(const Ctx& c, const Opnd& a, float* out, F f) {
    FOR_ELEM(k, e, t) out[e] = f(LDE(a, e)); END_ELEM
}

// ---- Math.h:9-28 ---------------------------------------------------------------------------------------
template <int NITER>
__device__ void op_unary(const Ctx& c, uint32_t mode, const Opnd& a, float* out) {
    switch (mode) {
        case U_SIN:   map1<NITER>(c, a, out, [](float x) { return sinf(x); }); break;
        case U_COS:   map1<NITER>(c, a, out, [](float x) { return cosf(x); }); break;
        case U_TAN:   map1<NITER>(c, a, out, [](float x) { return tanf(x); }); break;
        case U_TANH:  map1<NITER>(c, a, out, [](float x) { return tanhf(x); }); break;
        case U_ASINH: map1<NITER>(c, a, out, [](float x) { return asinhf(x); }); break;
        case U_LN:    map1<NITER>(c, a, out, [](float x) { return logf(x); }); break;
        case U_LOG10: map1<NITER>(c, a, out, [](float x) { return log10f(x); }); break;
        case U_LOG2:  map1<NITER>(c, a, out, [](float x) { return log2f(x); }); break;
        case U_CEIL:  map1<NITER>(c, a, out, [](float x) { return ceilf(x); }); break;
        case U_FLOOR: map1<NITER>(c, a, out, [](float x) { return floorf(x); }); break;
        case U_ROUND: map1<NITER>(c, a, out, [](float x) { return roundf(x); }); break;
        case U_SQRT:  map1<NITER>(c, a, out, [](float x) { return sqrtf(x); }); break;
        case U_EXP:   map1<NITER>(c, a, out, [](float x) { return expf(x); }); break;
        default:      map1<NITER>(c, a, out, [](float x) { return fabsf(x); }); break;
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

template <int NITER>
__device__ void op_reduce(const Ctx& c, uint32_t mode, const uint32_t* opnds, int n, float* out) {
    float acc[NITER];
    {
        const Opnd a = decode(c, __ldg(opnds));
        FOR_ELEM(k, e, t) acc[k] = LDE(a, e); END_ELEM
    }
    for (int j = 1; j < n; ++j) {
        const Opnd b = decode(c, __ldg(opnds + j));
        switch (mode) {   // hoisted so the inner loop is branch-free
            case R_ADD: FOR_ELEM(k, e, t) acc[k] = acc[k] + LDE(b, e); END_ELEM break;
            case R_SUB: FOR_ELEM(k, e, t) acc[k] = acc[k] - LDE(b, e); END_ELEM break;
            case R_MUL: FOR_ELEM(k, e, t) acc[k] = acc[k] * LDE(b, e); END_ELEM break;
            default:    FOR_ELEM(k, e, t) acc[k] = reduce_fn(mode, acc[k], LDE(b, e)); END_ELEM break;
        }
    }
    FOR_ELEM(k, e, t) out[e] = acc[k]; END_ELEM
}

__device__ __forceinline__ double bits_to_double(uint32_t lo, uint32_t hi) {
    return __longlong_as_double((long long) ((uint64_t) lo | ((uint64_t) hi << 32)));
}

} // namespace

// =========================================================================================================
template <int NITER>
__global__ void __launch_bounds__(128, 4) render_block_kernel(const __grid_constant__ LaunchParams P) {
    extern __shared__ __align__(16) float smem[];

    const int warpsPerCta = blockDim.x >> 5;
    const int warpInCta = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int L = P.tileWidth;
    const int logL = 31 - __clz(L);
    const int E = 32 * NITER;
    const int T = E >> logL;
    const int logT = 31 - __clz(T);
    const int per = 32 >> logL;           // samples of one voice held by one k-slice of the lanes
    const int nTiles = (P.nv + L - 1) / L;
    const int tile = blockIdx.x * warpsPerCta + warpInCta;
    if (tile >= nTiles) return;           // whole warp leaves together

    Ctx c;
    c.P = &P;
    c.lane = lane;
    c.vlane = lane & (L - 1);
    c.L = L; c.logL = logL; c.E = E; c.T = T;
    c.tile = tile;
    c.voice = tile * L + c.vlane;
    c.valid = c.voice < P.nv;
    c.owner = lane < L;
    const int perWarp = ((P.nSlots + P.nOut) * E + P.nStateRows * L + 3) & ~3;
    c.slots = smem + (size_t) warpInCta * perWarp;
    c.outacc = c.slots + P.nSlots * E;
    c.sst = c.outacc + P.nOut * E;

    // ---- state rows HBM -> shared memory (once per block), by the owner lanes ----
    if (c.owner) {
        int srow = 0;
        for (int i = 0; i < P.nStateEntries; ++i) {
            const uint32_t m = __ldg(P.stateMap + i);
            if (m == STATE_PAD) { srow += 1; continue; }
            const size_t row = m & ~STATE_DOUBLE_FLAG;
            if (m & STATE_DOUBLE_FLAG) {
                const double* g = reinterpret_cast<const double*>(P.rows + row * P.Vpad);
                reinterpret_cast<double*>(c.sst + srow * L)[lane] = g[c.voice];
                srow += 2;
            } else {
                c.sst[srow * L + lane] = P.rows[row * P.Vpad + c.voice];
                srow += 1;
            }
        }
    }
    __syncwarp();

    const int numSamples = P.numSamples;
    for (int s0 = 0; s0 < numSamples; s0 += T) {
        c.s0 = s0;
        c.cnt = min(T, numSamples - s0);

        for (int i = lane; i < P.nOut * E; i += 32) c.outacc[i] = 0.0f;

        const uint32_t* pc = P.code;
        for (;;) {
            __syncwarp();   // slot / state traffic of the previous op is visible to every lane
            const uint32_t w0 = __ldg(pc);
            const uint32_t opcode = w0 & 0xFF, nopnd = (w0 >> 8) & 0xFF, mode = w0 >> 24;
            if (opcode == OP_END) break;
            float* out = c.slots + ((w0 >> 16) & 0xFF) * E;
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
                FOR_ELEM(k, e, t) out[e] = 0.0f; END_ELEM
                break;

            case OP_COPY: {
                const Opnd a = decode(c, __ldg(opnds));
                FOR_ELEM(k, e, t) out[e] = LDE(a, e); END_ELEM
            } break;

            case OP_LOADIN: {
                if (P.inVoice) {
                    // transposed element order so that the lanes of a warp read consecutive samples of a voice
                    _Pragma("unroll") for (int k = 0; k < NITER; ++k) {
                        const int q = lane + 32 * k;
                        const int v = q >> logT, t = q & (T - 1);
                        const int voice = tile * L + v;
                        if (t < c.cnt) {
                            float x = 0.0f;
                            if (voice < P.nv) x = P.inVoice[((size_t) (P.voice0 + voice) * P.nIn + aux0) * P.inStride + s0 + t];
                            out[t * L + v] = x;
                        }
                    }
                } else {
                    const float* g = P.inShared + (size_t) aux0 * P.inStride + s0;
                    FOR_ELEM(k, e, t) out[e] = __ldg(g + t); END_ELEM
                }
            } break;

            case OP_UNARY: {
                const Opnd a = decode(c, __ldg(opnds));
                op_unary<NITER>(c, mode, a, out);
            } break;

            case OP_BINARY: {
                const Opnd a = decode(c, __ldg(opnds));
                const Opnd b = decode(c, __ldg(opnds + 1));
                FOR_ELEM(k, e, t) out[e] = binary_fn(mode, LDE(a, e), LDE(b, e)); END_ELEM
            } break;

            case OP_REDUCE:
                op_reduce<NITER>(c, mode, opnds, (int) nopnd, out);
                break;

            case OP_PHASOR: {   // Core.h:89-97: step = f * (1/sr) in float; phase = next - floor(next)
                const Opnd f = decode(c, __ldg(opnds));
                const float rsr = __uint_as_float(aux0);
                if (c.owner) {
                    float phase = c.sst[sidx * L + lane];
                    FOR_OWNER(t) {
                        const float step = LDT(f, t) * rsr;
                        out[t * L + lane] = phase;
                        const float next = phase + step;
                        phase = next - floorf(next);
                    }
                    c.sst[sidx * L + lane] = phase;
                }
            } break;

            case OP_SPHASOR: {  // Core.h:113-121; state: phase, change.lastIn
                const Opnd f = decode(c, __ldg(opnds));
                const Opnd r = decode(c, __ldg(opnds + 1));
                const float rsr = __uint_as_float(aux0);
                if (c.owner) {
                    float phase = c.sst[sidx * L + lane];
                    float last = c.sst[(sidx + 1) * L + lane];
                    FOR_OWNER(t) {
                        const float xn = LDT(f, t);
                        if (change_tick(last, LDT(r, t)) > 0.5f) phase = 0.0f;
                        const float step = xn * rsr;
                        out[t * L + lane] = phase;
                        const float next = phase + step;
                        phase = next - floorf(next);
                    }
                    c.sst[sidx * L + lane] = phase;
                    c.sst[(sidx + 1) * L + lane] = last;
                }
            } break;

            case OP_COUNTER: {  // Core.h:198-211
                const Opnd g = decode(c, __ldg(opnds));
                if (c.owner) {
                    float count = c.sst[sidx * L + lane];
                    FOR_OWNER(t) {
                        const float in = LDT(g, t);
                        if ((1.0f - in) <= kEps) { out[t * L + lane] = count; count = count + 1.0f; }
                        else { count = 0.0f; out[t * L + lane] = 0.0f; }
                    }
                    c.sst[sidx * L + lane] = count;
                }
            } break;

            case OP_ACCUM: {    // Core.h:233-243; state: runningTotal, change.lastIn
                const Opnd x = decode(c, __ldg(opnds));
                const Opnd r = decode(c, __ldg(opnds + 1));
                if (c.owner) {
                    float total = c.sst[sidx * L + lane];
                    float last = c.sst[(sidx + 1) * L + lane];
                    FOR_OWNER(t) {
                        if (change_tick(last, LDT(r, t)) > 0.5f) total = 0.0f;
                        total += LDT(x, t);
                        out[t * L + lane] = total;
                    }
                    c.sst[sidx * L + lane] = total;
                    c.sst[(sidx + 1) * L + lane] = last;
                }
            } break;

            case OP_LATCH: {    // Core.h:265-281; state: z, hold
                const Opnd l = decode(c, __ldg(opnds));
                const Opnd x = decode(c, __ldg(opnds + 1));
                if (c.owner) {
                    float z = c.sst[sidx * L + lane];
                    float hold = c.sst[(sidx + 1) * L + lane];
                    FOR_OWNER(t) {
                        const float lv = LDT(l, t);
                        if (fabsf(z) <= kEps && lv > kEps) hold = LDT(x, t);
                        z = lv;
                        out[t * L + lane] = hold;
                    }
                    c.sst[sidx * L + lane] = z;
                    c.sst[(sidx + 1) * L + lane] = hold;
                }
            } break;

            case OP_MAXHOLD: {  // Core.h:315-332; state: max, samplesAtCurrentMax(u32), change.lastIn; aux0 = holdTimeSamples
                const Opnd x = decode(c, __ldg(opnds));
                const Opnd r = decode(c, __ldg(opnds + 1));
                if (c.owner) {
                    float mx = c.sst[sidx * L + lane];
                    uint32_t held = __float_as_uint(c.sst[(sidx + 1) * L + lane]);
                    float last = c.sst[(sidx + 2) * L + lane];
                    const uint32_t hts = aux0;
                    FOR_OWNER(t) {
                        const float in = LDT(x, t);
                        bool reset = change_tick(last, LDT(r, t)) > 0.5f;
                        if (!reset) reset = (++held >= hts);   // short-circuit || of the reference
                        if (reset) { mx = in; held = 0; }
                        else if (in > mx) { held = 0; mx = in; }
                        out[t * L + lane] = mx;
                    }
                    c.sst[sidx * L + lane] = mx;
                    c.sst[(sidx + 1) * L + lane] = __uint_as_float(held);
                    c.sst[(sidx + 2) * L + lane] = last;
                }
            } break;

            case OP_RAND: {     // Noise.h:25-38
                if (c.owner) {
                    uint32_t seed = __float_as_uint(c.sst[sidx * L + lane]);
                    FOR_OWNER(t) {
                        seed = 214013u * seed + 2531011u;
                        const int r = (int) ((seed >> 16) & 0x7FFFu);
                        out[t * L + lane] = (float) r / 32767.0f;
                    }
                    c.sst[sidx * L + lane] = __uint_as_float(seed);
                }
            } break;

            case OP_POLE: {     // Filters.h:27-33
                const Opnd pp = decode(c, __ldg(opnds));
                const Opnd x = decode(c, __ldg(opnds + 1));
                if (c.owner) {
                    float z = c.sst[sidx * L + lane];
                    FOR_OWNER(t) {
                        z = LDT(x, t) + LDT(pp, t) * z;
                        out[t * L + lane] = z;
                    }
                    c.sst[sidx * L + lane] = z;
                }
            } break;

            case OP_ENV: {      // Filters.h:61-73
                const Opnd ap = decode(c, __ldg(opnds));
                const Opnd rp = decode(c, __ldg(opnds + 1));
                const Opnd x = decode(c, __ldg(opnds + 2));
                if (c.owner) {
                    float z = c.sst[sidx * L + lane];
                    FOR_OWNER(t) {
                        const float vn = fabsf(LDT(x, t));
                        const float pcoef = (vn > z) ? LDT(ap, t) : LDT(rp, t);
                        z = pcoef * (z - vn) + vn;
                        out[t * L + lane] = z;
                    }
                    c.sst[sidx * L + lane] = z;
                }
            } break;

            case OP_BIQUAD: {   // Filters.h:102-114 (TDF-II, audio-rate coefficients)
                const Opnd b0 = decode(c, __ldg(opnds));
                const Opnd b1 = decode(c, __ldg(opnds + 1));
                const Opnd b2 = decode(c, __ldg(opnds + 2));
                const Opnd a1 = decode(c, __ldg(opnds + 3));
                const Opnd a2 = decode(c, __ldg(opnds + 4));
                const Opnd x = decode(c, __ldg(opnds + 5));
                if (c.owner) {
                    float z1 = c.sst[sidx * L + lane];
                    float z2 = c.sst[(sidx + 1) * L + lane];
                    FOR_OWNER(t) {
                        const float xn = LDT(x, t);
                        const float y = LDT(b0, t) * xn + z1;
                        z1 = LDT(b1, t) * xn - LDT(a1, t) * y + z2;
                        z2 = LDT(b2, t) * xn - LDT(a2, t) * y;
                        out[t * L + lane] = y;
                    }
                    c.sst[sidx * L + lane] = z1;
                    c.sst[(sidx + 1) * L + lane] = z2;
                }
            } break;

            case OP_PREWARP: {  // filters/MultiMode1p.h:23-33; (aux0,aux1) = bits of T = 1.0/sr (double)
                const Opnd fc = decode(c, __ldg(opnds));
                const double Ts = bits_to_double(aux0, aux1);
                FOR_ELEM(k, e, t) {
                    const double twoPi = 2.0 * 3.141592653589793238;
                    const double wd = twoPi * (double) LDE(fc, e);
                    out[e] = (float) tan(wd * Ts / 2.0);
                } END_ELEM
            } break;

            case OP_MM1P: {     // filters/MultiMode1p.h:78-103; state: double z; mode 0 low / 2 high / 4 all
                const Opnd gi = decode(c, __ldg(opnds));
                const Opnd x = decode(c, __ldg(opnds + 1));
                if (c.owner) {
                    double* zs = reinterpret_cast<double*>(c.sst + sidx * L) + lane;
                    double z = *zs;
                    FOR_OWNER(t) {
                        const double g = clampd((double) LDT(gi, t), 0.0, 0.9999);
                        const float xn = LDT(x, t);
                        const double G = g / (1.0 + g);
                        const double v = ((double) xn - z) * G;
                        const double lp = v + z;
                        z = lp + v;
                        float y;
                        if (mode == 0) y = (float) lp;
                        else if (mode == 2) y = xn - (float) lp;
                        else y = (float) (lp + lp - (double) xn);
                        out[t * L + lane] = y;
                    }
                    *zs = z;
                }
            } break;

            case OP_SVF: {      // filters/SVF.h:48-104; (aux0,aux1) = bits of sr (double); state: double ic1eq, ic2eq
                const Opnd fc = decode(c, __ldg(opnds));
                const Opnd q = decode(c, __ldg(opnds + 1));
                const Opnd x = decode(c, __ldg(opnds + 2));
                const double sr = bits_to_double(aux0, aux1);
                const double fmax = sr / 2.0001;
                // phase 1 — coefficients (updateCoeffs, SVF.h:72-80) are a pure function of (fc, q) per sample:
                // all lanes, one element each per k
                double ga[NITER], a1a[NITER], ka[NITER];
                _Pragma("unroll") for (int k = 0; k < NITER; ++k) {
                    const int e = lane + 32 * k;
                    const double g = tan(3.14159265359 * clampd((double) LDE(fc, e), 20.0, fmax) / sr);
                    const double kq = 1.0 / clampd((double) LDE(q, e), 0.25, 20.0);
                    ga[k] = g; ka[k] = kq;
                    a1a[k] = 1.0 / (1.0 + g * (g + kq));
                }
                // phase 2 — the tick recurrence (SVF.h:48-70), serial per voice; every lane walks the loop so the
                // coefficients can be fetched from the lane that computed them (sample t of voice v lives in
                // lane (t*L + v) & 31 of slice k = (t*L) >> 5)
                double ic1 = 0.0, ic2 = 0.0;
                if (c.owner) {
                    ic1 = reinterpret_cast<const double*>(c.sst + sidx * L)[lane];
                    ic2 = reinterpret_cast<const double*>(c.sst + (sidx + 2) * L)[lane];
                }
                _Pragma("unroll") for (int k = 0; k < NITER; ++k) {
                    for (int j = 0; j < per; ++j) {
                        const int t = k * per + j;
                        if (t >= c.cnt) break;     // warp-uniform
                        const int src = j * L + c.vlane;
                        const double g = __shfl_sync(FULL, ga[k], src);
                        const double a1 = __shfl_sync(FULL, a1a[k], src);
                        const double kq = __shfl_sync(FULL, ka[k], src);
                        if (c.owner) {
                            const double a2 = g * a1;
                            const double a3 = g * a2;
                            const float v0 = LDT(x, t);
                            const double v3 = (double) v0 - ic2;
                            const double v1 = ic1 * a1 + v3 * a2;
                            const double v2 = ic2 + ic1 * a2 + v3 * a3;
                            ic1 = v1 * 2.0 - ic1;
                            ic2 = v2 * 2.0 - ic2;
                            float y;
                            switch (mode) {
                                case 0: y = (float) v2; break;
                                case 1: y = (float) v1; break;
                                case 2: y = (float) ((double) v0 - kq * v1 - v2); break;
                                case 3: y = (float) ((double) v0 - kq * v1); break;
                                default: y = (float) ((double) v0 - 2.0 * kq * v1); break;
                            }
                            out[t * L + lane] = y;
                        }
                    }
                }
                if (c.owner) {
                    reinterpret_cast<double*>(c.sst + sidx * L)[lane] = ic1;
                    reinterpret_cast<double*>(c.sst + (sidx + 2) * L)[lane] = ic2;
                }
            } break;

            case OP_SVFSHELF: { // filters/SVFShelf.h:44-106; mode 0 lowshelf / 1 highshelf / 2 bell
                const Opnd fc = decode(c, __ldg(opnds));
                const Opnd q = decode(c, __ldg(opnds + 1));
                const Opnd gdb = decode(c, __ldg(opnds + 2));
                const Opnd x = decode(c, __ldg(opnds + 3));
                const double sr = bits_to_double(aux0, aux1);
                const double fmax = sr / 2.0001;
                if (c.owner) {
                    double* s1 = reinterpret_cast<double*>(c.sst + sidx * L) + lane;
                    double* s2 = reinterpret_cast<double*>(c.sst + (sidx + 2) * L) + lane;
                    double ic1 = *s1, ic2 = *s2;
                    _Pragma("unroll 2") for (int t = 0; t < c.cnt; ++t) {
                        const double A = pow(10.0, (double) LDT(gdb, t) / 40.0);
                        double g = tan(3.14159265359 * clampd((double) LDT(fc, t), 20.0, fmax) / sr);
                        double kq = 1.0 / clampd((double) LDT(q, t), 0.25, 20.0);
                        if (mode == 0) g /= A;
                        if (mode == 1) g *= A;
                        if (mode == 2) kq /= A;
                        const double a1 = 1.0 / (1.0 + g * (g + kq));
                        const double a2 = g * a1;
                        const double a3 = g * a2;
                        const float v0 = LDT(x, t);
                        const double v3 = (double) v0 - ic2;
                        const double v1 = ic1 * a1 + v3 * a2;
                        const double v2 = ic2 + ic1 * a2 + v3 * a3;
                        ic1 = v1 * 2.0 - ic1;
                        ic2 = v2 * 2.0 - ic2;
                        float y;
                        if (mode == 2) y = (float) ((double) v0 + kq * (A * A - 1.0) * v1);
                        else if (mode == 0) y = (float) ((double) v0 + kq * (A - 1.0) * v1 + (A * A - 1.0) * v2);
                        else y = (float) (A * A * (double) v0 + kq * (1.0 - A) * A * v1 + (1.0 - A * A) * v2);
                        out[t * L + lane] = y;
                    }
                    *s1 = ic1; *s2 = ic2;
                }
            } break;

            case OP_Z: {        // Delays.h:29-34 — out[t] = in[t-1]: no recurrence, only a carry between tiles
                const Opnd x = decode(c, __ldg(opnds));
                const float zprev = c.sst[sidx * L + c.vlane];
                FOR_ELEM(k, e, t) out[e] = (t == 0) ? zprev : LDE(x, e - L); END_ELEM
                __syncwarp();
                if (c.owner) c.sst[sidx * L + lane] = LDT(x, c.cnt - 1);
            } break;

            case OP_DELAY: {    // Delays.h:108-159; aux0 = size; ring [tile][pos][L]; state: writeIndex
                const Opnd len = decode(c, __ldg(opnds));
                const Opnd fb = decode(c, __ldg(opnds + 1));
                const Opnd x = decode(c, __ldg(opnds + 2));
                const int size = (int) aux0;
                if (size == 0) { FOR_ELEM(k, e, t) out[e] = LDE(len, e); END_ELEM break; }   // Delays.h:105-106 copies inputData[0]
                float* ring = reinterpret_cast<float*>(ptrbits) + (size_t) tile * size * L + c.vlane;
                const float fsize = (float) size;
                const int w0 = __float_as_int(c.sst[sidx * L + c.vlane]);
                // Fast path: when no read head of this tile can land on a position written inside the tile (and
                // no write of the tile can hit a position still to be read), samples are independent.
                bool hazard = false;
                FOR_ELEM(k, e, t) {
                    const float offset = clampf(LDE(len, e), 0.0f, fsize);
                    if (!(offset <= kEps) && !(offset >= (float) (c.cnt + 1) && offset <= (float) (size - c.cnt - 1))) hazard = true;
                } END_ELEM
                if (!__any_sync(FULL, hazard)) {
                    FOR_ELEM(k, e, t) {
                        int w = w0 + t;
                        while (w >= size) w -= size;
                        const float offset = clampf(LDE(len, e), 0.0f, fsize);
                        float y, in;
                        if (offset <= kEps) { in = LDE(x, e); y = in; }
                        else {
                            const float readFrac = (float) (size + w) - offset;
                            int readLeft = (int) readFrac;
                            const float frac = readFrac - floorf(readFrac);
                            int readRight = readLeft + 1;                       // both in [0, 2*size]: % size by subtraction
                            if (readLeft >= size) readLeft -= size;
                            if (readRight >= size) readRight -= size;
                            if (readRight >= size) readRight -= size;
                            const float left = ring[(size_t) readLeft * L];
                            const float right = ring[(size_t) readRight * L];
                            y = left + frac * (right - left);
                            in = LDE(x, e) + clampf(LDE(fb, e), -1.0f, 1.0f) * y;
                        }
                        ring[(size_t) w * L] = in;
                        out[e] = y;
                    } END_ELEM
                    __syncwarp();
                    if (c.owner) {
                        int w = w0 + c.cnt;
                        while (w >= size) w -= size;
                        c.sst[sidx * L + lane] = __int_as_float(w);
                    }
                } else if (c.owner) {
                    int w = w0;
                    FOR_OWNER(t) {
                        const float offset = clampf(LDT(len, t), 0.0f, fsize);
                        float y, in;
                        if (offset <= kEps) { in = LDT(x, t); y = in; }
                        else {
                            const float readFrac = (float) (size + w) - offset;
                            const int readLeft = (int) readFrac;
                            const int readRight = readLeft + 1;
                            const float frac = readFrac - floorf(readFrac);
                            const float left = ring[(size_t) (readLeft % size) * L];
                            const float right = ring[(size_t) (readRight % size) * L];
                            y = left + frac * (right - left);
                            in = LDT(x, t) + clampf(LDT(fb, t), -1.0f, 1.0f) * y;
                        }
                        ring[(size_t) w * L] = in;
                        out[t * L + lane] = y;
                        if (++w >= size) w -= size;
                    }
                    c.sst[sidx * L + lane] = __int_as_float(w);
                }
            } break;

            case OP_SDELAY: {   // Delays.h:246-260; aux0 = ring size (pow2), aux1 = length; state: writeIndex
                // out[t] = the sample written `len` samples ago: inside the tile it is still in the input slot,
                // older ones are in the ring (size >= len + blockSize keeps reads and writes of a tile disjoint).
                const Opnd x = decode(c, __ldg(opnds));
                const int size = (int) aux0, mask = size - 1, len = (int) aux1;
                float* ring = reinterpret_cast<float*>(ptrbits) + (size_t) tile * size * L + c.vlane;
                const int w0 = __float_as_int(c.sst[sidx * L + c.vlane]);
                FOR_ELEM(k, e, t) {
                    float y;
                    if (t >= len) y = LDE(x, e - len * L);
                    else y = ring[(size_t) ((w0 + t - len + size) & mask) * L];
                    out[e] = y;
                } END_ELEM
                FOR_ELEM(k, e, t) ring[(size_t) ((w0 + t) & mask) * L] = LDE(x, e); END_ELEM
                __syncwarp();
                if (c.owner) c.sst[sidx * L + lane] = __int_as_float((w0 + c.cnt) & mask);
            } break;

            case OP_TABLE: {    // Table.h:59-71; aux0 = table length; ptr = device copy of resource channel 0
                const Opnd pos = decode(c, __ldg(opnds));
                const int size = (int) aux0;
                const float* tab = reinterpret_cast<const float*>(ptrbits);
                FOR_ELEM(k, e, t) {
                    const float readPos = clampf(LDE(pos, e), 0.0f, 1.0f) * (float) (size - 1);
                    const int readLeft = (int) readPos;
                    const int readRight = readLeft + 1;
                    const float frac = readPos - floorf(readPos);
                    const float left = __ldg(tab + (readLeft % size));
                    const float right = __ldg(tab + (readRight % size));
                    out[e] = left + frac * (right - left);
                } END_ELEM
            } break;

            case OP_BLEP: {     // Oscillators.h:23-89; state: phase, acc; aux0 = bits of float(sr)
                const Opnd f = decode(c, __ldg(opnds));
                const float sr = __uint_as_float(aux0);
                auto blep = [](float ph, float inc) -> float {
                    if (ph < inc) { const float p = ph / inc; return (2.0f - p) * p - 1.0f; }
                    if (ph > (1.0f - inc)) { const float p = (ph - 1.0f) / inc; return (p + 2.0f) * p + 1.0f; }
                    return 0.0f;
                };
                if (c.owner) {
                    float phase = c.sst[sidx * L + lane];
                    float acc = c.sst[(sidx + 1) * L + lane];
                    FOR_OWNER(t) {
                        const float inc = LDT(f, t) / sr;
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
                        out[t * L + lane] = y;
                    }
                    c.sst[sidx * L + lane] = phase;
                    c.sst[(sidx + 1) * L + lane] = acc;
                }
            } break;

            case OP_TAPIN: {    // Feedback.h:42-52; ptr = shared tap buffer [tile][blockSize][L] (element order)
                const float* tap = reinterpret_cast<const float*>(ptrbits) + ((size_t) tile * P.blockSize + s0) * L;
                FOR_ELEM(k, e, t) out[e] = tap[e]; END_ELEM
            } break;

            case OP_TAPOUT: {   // Feedback.h:109-121; ptr = this node's private delayBuffer [tile][blockSize][L]
                const Opnd x = decode(c, __ldg(opnds));
                float* buf = reinterpret_cast<float*>(ptrbits) + ((size_t) tile * P.blockSize + s0) * L;
                FOR_ELEM(k, e, t) { const float v = LDE(x, e); buf[e] = v; out[e] = v; } END_ELEM
            } break;

            case OP_STOREBUF: { // stage boundary: ptr = [voice][blockSize] staging buffer (transposed element order)
                const Opnd x = decode(c, __ldg(opnds));
                float* base = reinterpret_cast<float*>(ptrbits);
                const uint32_t w = __ldg(opnds);
                _Pragma("unroll") for (int k = 0; k < NITER; ++k) {
                    const int qq = lane + 32 * k;
                    const int v = qq >> logT, t = qq & (T - 1);
                    const int voice = tile * L + v;
                    if (t < c.cnt && voice < P.nv) {
                        float val = 0.0f;
                        if (x.slot) val = x.p[t * L + v];
                        else if ((w >> 30) == K_PARAM) val = __ldg(P.rows + (size_t) (w & 0x3FFFFFFFu) * P.Vpad + voice);
                        base[(size_t) voice * P.blockSize + s0 + t] = val;
                    }
                }
            } break;

            case OP_LOADBUF: {
                const float* base = reinterpret_cast<const float*>(ptrbits);
                _Pragma("unroll") for (int k = 0; k < NITER; ++k) {
                    const int qq = lane + 32 * k;
                    const int v = qq >> logT, t = qq & (T - 1);
                    const int voice = tile * L + v;
                    if (t < c.cnt) out[t * L + v] = (voice < P.nv) ? base[(size_t) voice * P.blockSize + s0 + t] : 0.0f;
                }
            } break;

            case OP_ROOT: {     // Core.h:66-78 + GainFade.h:56-72; aux0 = root index
                const RootDyn rd = P.roots[aux0];
                if (nopnd < 1) { FOR_ELEM(k, e, t) out[e] = 0.0f; END_ELEM }
                else {
                    const Opnd x = decode(c, __ldg(opnds));
                    if (rd.gain0 == rd.target) {
                        FOR_ELEM(k, e, t) out[e] = LDE(x, e) * rd.target; END_ELEM
                    } else {
                        FOR_ELEM(k, e, t) {
                            const float g = clampf(rd.gain0 + rd.step * (float) (s0 + t), 0.0f, 1.0f);
                            out[e] = LDE(x, e) * g;
                        } END_ELEM
                    }
                }
                if (rd.channel >= 0 && rd.channel < P.nOut) {   // GraphRenderSequence.h:227-231
                    float* acc = c.outacc + rd.channel * E;
                    FOR_ELEM(k, e, t) acc[e] += out[e]; END_ELEM
                }
            } break;

            default: break;
            }
        }

        // ---- tile epilogue: per-voice output and per-tile partial mix ----
        if (P.outVoice) {
            for (int ch = 0; ch < P.nOut; ++ch) {
                const float* a = c.outacc + ch * E;
                _Pragma("unroll") for (int k = 0; k < NITER; ++k) {
                    const int qq = lane + 32 * k;           // transposed: consecutive lanes = consecutive samples
                    const int v = qq >> logT, t = qq & (T - 1);
                    const int voice = tile * L + v;
                    if (t < c.cnt && voice < P.nv)
                        P.outVoice[((size_t) (P.voice0 + voice) * P.nOut + ch) * P.outStride + s0 + t] = a[t * L + v];
                }
            }
        }
        if (P.mixPartial) {
            // sum over the voices of the tile in a fixed order (xor butterfly over the voice bits of the lane id)
            for (int ch = 0; ch < P.nOut; ++ch) {
                const float* a = c.outacc + ch * E;
                float* gp = P.mixPartial + ((size_t) (P.tileBase + tile) * P.nOut + ch) * P.blockSize + s0;
                _Pragma("unroll") for (int k = 0; k < NITER; ++k) {
                    const int e = lane + 32 * k;
                    const int t = e >> logL;
                    float v = (c.valid && t < c.cnt) ? a[e] : 0.0f;
                    for (int d = L >> 1; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
                    if (c.vlane == 0 && t < c.cnt) gp[t] = v;
                }
            }
        }
    }
    __syncwarp();

    // ---- state rows shared memory -> HBM ----
    if (c.owner) {
        int srow = 0;
        for (int i = 0; i < P.nStateEntries; ++i) {
            const uint32_t m = __ldg(P.stateMap + i);
            if (m == STATE_PAD) { srow += 1; continue; }
            const size_t row = m & ~STATE_DOUBLE_FLAG;
            if (m & STATE_DOUBLE_FLAG) {
                double* g = reinterpret_cast<double*>(P.rows + row * P.Vpad);
                g[c.voice] = reinterpret_cast<const double*>(c.sst + srow * L)[lane];
                srow += 2;
            } else {
                P.rows[row * P.Vpad + c.voice] = c.sst[srow * L + lane];
                srow += 1;
            }
        }
    }

    // ---- tap promotion (GraphRenderSequence.h:200-210,306-308): records after OP_END, until the second OP_END ----
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
            const float* src = reinterpret_cast<const float*>(sb) + (size_t) tile * P.blockSize * L;
            float* dst = reinterpret_cast<float*>(db) + (size_t) tile * P.blockSize * L;
            for (int i = lane; i < numSamples * L; i += 32) dst[i] = src[i];
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

int render_niter_for(int tileWidth) {
    // E = L*T elements per sample tile: 256 when the tile is wide enough (T = 256/L >= 8), else T = 32 samples.
    if (tileWidth >= 8) return 8;
    return tileWidth;   // L = 4 -> 4, 2 -> 2, 1 -> 1
}

size_t render_smem_bytes(int nSlots, int nOut, int nStateRows, int warpsPerCta, int tileWidth) {
    const int E = 32 * render_niter_for(tileWidth);
    const size_t perWarp = ((size_t) (nSlots + nOut) * E + (size_t) nStateRows * tileWidth + 3) & ~(size_t) 3;
    return (size_t) warpsPerCta * perWarp * sizeof(float);
}

template <int NITER>
static cudaError_t launch_impl(const LaunchParams& P, int grid, int threads, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(render_block_kernel<NITER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    render_block_kernel<NITER><<<grid, threads, smem, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_render_block(const LaunchParams& P, int warpsPerCta, cudaStream_t stream) {
    const int L = P.tileWidth;
    const int nTiles = (P.nv + L - 1) / L;
    if (nTiles <= 0) return cudaSuccess;
    const int grid = (nTiles + warpsPerCta - 1) / warpsPerCta;
    const size_t smem = render_smem_bytes(P.nSlots, P.nOut, P.nStateRows, warpsPerCta, L);
    switch (render_niter_for(L)) {
        case 8: return launch_impl<8>(P, grid, warpsPerCta * 32, smem, stream);
        case 4: return launch_impl<4>(P, grid, warpsPerCta * 32, smem, stream);
        case 2: return launch_impl<2>(P, grid, warpsPerCta * 32, smem, stream);
        default: return launch_impl<1>(P, grid, warpsPerCta * 32, smem, stream);
    }
}

cudaError_t launch_mix_reduce(const float* partial, float* out, int nTiles, int nOut, int blockSize, int numSamples, cudaStream_t stream) {
    const int chunksPerCh = (blockSize + 31) / 32;
    mix_reduce_kernel<<<nOut * chunksPerCh, 1024, 0, stream>>>(partial, out, nTiles, nOut, blockSize, numSamples);
    return cudaGetLastError();
}

} // namespace eb
