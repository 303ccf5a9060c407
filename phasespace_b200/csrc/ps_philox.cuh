// Philox4x32-10 counter-based RNG (Salmon et al., SC'11), keyed on (seed, event index).
//
// Replaces the reference's tf.random.Generator (random.py:14-40; draw sites phasespace.py:367,
// 444-446, 451).  One call yields two fp64 uniforms.  Stream layout (must match oracle/philox.py):
//   key = (seed_lo, seed_hi); counter = (event_lo, event_hi, column/2, stream)
//   uniform(column) = ((out[2k+1] << 32 | out[2k]) >> 12) * 2^-52,  k = column % 2
//   (built as the 52-bit mantissa of a double in [1,2) minus 1: one shift/or and one DADD)
// so a draw is a pure function of (seed, global event index, column): any sharding of the event
// range over GPUs, launches or tiles gives bit-identical events.
#ifndef PS_PHILOX_CUH_
#define PS_PHILOX_CUH_

// 1: inline mul.wide.u32; 0: __umulhi + mul (the compiler fuses them into IMAD.WIDE as well and keeps the
// round keys in uniform registers, which costs fewer vector registers)
#ifndef PS_PHILOX_ASM
#define PS_PHILOX_ASM 0
#endif

namespace psk {

struct Philox2d {
    double d0, d1;
};

__device__ __forceinline__ Philox2d philox_pair(unsigned long long seed, unsigned long long event,
                                               unsigned int block, unsigned int stream) {
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
    Philox2d out;
    out.d0 = __hiloint2double((int)(0x3FF00000u | (c1 >> 12)), (int)((c1 << 20) | (c0 >> 12))) - 1.0;
    out.d1 = __hiloint2double((int)(0x3FF00000u | (c3 >> 12)), (int)((c3 << 20) | (c2 >> 12))) - 1.0;
    return out;
}

}  // namespace psk
#endif
