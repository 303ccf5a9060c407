// Philox4x32-10 counter-based RNG (Salmon et al., SC'11), keyed on (seed, event index).
//
// Replaces the reference's tf.random.Generator (random.py:14-40; draw sites phasespace.py:367,
// 444-446, 451).  Stream layout (must match oracle/philox.py):
//   key = (seed_lo, seed_hi); counter = (event_lo, event_hi, block, stream)
//   the 128-bit outputs of blocks 0, 1, 2, ... of one event form a bit string (word j of block b holds
//   bits 128 b + 32 j ... + 31, least significant first); uniform number `column` is the 51-bit field at
//   bit offset 51 * column, scaled by 2^-51.
// Dense 51-bit packing matters: the headline three-body decay needs 5 uniforms = 255 bits = two Philox
// blocks instead of three.  A draw is a pure function of (seed, global event index, column): any sharding
// of the event range over GPUs, launches or tiles gives bit-identical events.
#ifndef PS_PHILOX_CUH_
#define PS_PHILOX_CUH_

// 1: inline mul.wide.u32; 0: __umulhi + mul (the compiler fuses them into IMAD.WIDE as well and keeps the
// round keys in uniform registers, which costs fewer vector registers)
#ifndef PS_PHILOX_ASM
#define PS_PHILOX_ASM 0
#endif

#define PS_UNIFORM_BITS 51

namespace psk {

struct Philox4 {
    unsigned int w[4];
};

__device__ __forceinline__ Philox4 philox4x32_10(unsigned long long seed, unsigned long long event, unsigned int block,
                                                 unsigned int stream) {
    unsigned int c0 = (unsigned int)event, c1 = (unsigned int)(event >> 32), c2 = block, c3 = stream;
    unsigned int k0 = (unsigned int)seed, k1 = (unsigned int)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        // one IMAD.WIDE.U32 per multiplier yields both halves of the 32x32 product
        unsigned int hi0, lo0, hi1, lo1;
#if PS_PHILOX_ASM
        asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}" : "=r"(lo0), "=r"(hi0) : "r"(c0), "r"(0xD2511F53u));
        asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}" : "=r"(lo1), "=r"(hi1) : "r"(c2), "r"(0xCD9E8D57u));
#else
        hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#endif
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 out;
    out.w[0] = c0;
    out.w[1] = c1;
    out.w[2] = c2;
    out.w[3] = c3;
    return out;
}

// The same with the ten round keys precomputed (rk[2 r], rk[2 r + 1] = key words of round r: the host derives them from
// the seed into the kernel parameter block, PsParams::philox_keys), so that the kernel reads them as constants instead of
// re-deriving them with twenty uniform-datapath additions per tile.
__device__ __forceinline__ Philox4 philox4x32_10_keyed(const unsigned int* rk, unsigned long long event, unsigned int block,
                                                       unsigned int stream) {
    unsigned int c0 = (unsigned int)event, c1 = (unsigned int)(event >> 32), c2 = block, c3 = stream;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ rk[2 * r];
        c1 = lo1;
        c2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c3 = lo0;
    }
    Philox4 out;
    out.w[0] = c0;
    out.w[1] = c1;
    out.w[2] = c2;
    out.w[3] = c3;
    return out;
}

// 51 random bits (lo: 32, hi: low 19 bits used) -> U[0,1) with resolution 2^-51: the field fills the low
// 51 mantissa bits of a double d in [1, 1.5), and u = 2 d - 2 is exact.
__device__ __forceinline__ double uniform_from_bits(unsigned int lo, unsigned int hi) {
    return fma(__hiloint2double((int)((hi & 0x7FFFFu) | 0x3FF00000u), (int)lo), 2.0, -2.0);
}

// The two words of that double d for the field at bit offset O (0..127) of the bit string x[0..7] (two consecutive
// Philox blocks, or one block and zeros): lo = bits [O, O+32); hiword = exponent 0x3FF over bits [O+32, O+51).
// (Sliding the 19 high bits under the exponent with a third funnel shift, __funnelshift_r(t, 0x7FE, 13), saves the
// or, but measured no faster: the constant then occupies a register.)
template <int O>
__device__ __forceinline__ void field_words(const unsigned int* x, unsigned int& lo, unsigned int& hiword) {
    constexpr int i = O / 32, sh = O % 32, j = (O + 19) / 32, s2 = (O + 19) % 32;
    lo = sh ? __funnelshift_r(x[i], x[i + 1], sh) : x[i];
    const unsigned int t = s2 ? __funnelshift_r(x[j], x[j + 1], s2) : x[j];  // the 19 high bits on top of a word
    hiword = (t >> 13) | 0x3FF00000u;
}

// uniform number COL of an event in any stream (same bit layout as Ctx::next in ps_kernel.cuh, blocks recomputed
// per call: for the small helper kernels that draw two or three numbers per event)
template <int COL>
__device__ __forceinline__ double philox_uniform_col(unsigned long long seed, unsigned long long event, unsigned int stream) {
    constexpr int bit0 = PS_UNIFORM_BITS * COL;
    constexpr int FB = bit0 / 128, LB = (bit0 + PS_UNIFORM_BITS - 1) / 128;
    const Philox4 a = philox4x32_10(seed, event, FB, stream);
    unsigned int x[8] = {a.w[0], a.w[1], a.w[2], a.w[3], 0u, 0u, 0u, 0u};
    if constexpr (LB > FB) {
        const Philox4 b = philox4x32_10(seed, event, LB, stream);
        x[4] = b.w[0], x[5] = b.w[1], x[6] = b.w[2], x[7] = b.w[3];
    }
    unsigned int lo, hiword;
    field_words<bit0 - 128 * FB>(x, lo, hiword);
    return fma(__hiloint2double((int)hiword, (int)lo), 2.0, -2.0);
}

// first uniform of block 0 of a stream (acceptance draws, stand-alone samplers)
__device__ __forceinline__ double philox_uniform0(unsigned long long seed, unsigned long long event, unsigned int stream) {
    const Philox4 r = philox4x32_10(seed, event, 0u, stream);
    return uniform_from_bits(r.w[0], r.w[1]);
}

__device__ __forceinline__ double philox_uniform0_keyed(const unsigned int* rk, unsigned long long event, unsigned int stream) {
    const Philox4 r = philox4x32_10_keyed(rk, event, 0u, stream);
    return uniform_from_bits(r.w[0], r.w[1]);
}

}  // namespace psk
#endif
