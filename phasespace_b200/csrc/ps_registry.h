/* Registry of ahead-of-time compiled kernels: each ps_aot.cu translation unit registers its table
 * with the launcher (ps_api.cu) from a static initialiser. */
#ifndef PS_REGISTRY_H_
#define PS_REGISTRY_H_

#define PS_VARIANT_GENERATE 0 /* genbod_kernel: weights + 4-vectors */
#define PS_VARIANT_MASK 1     /* genbod_mask_kernel: weights only -> accept mask (unweighting pass 1) */
#define PS_VARIANT_HIST 2     /* genbod_hist_kernel: fused weighted histograms, nothing stored */

struct PsAotEntry {
    const char* signature; /* canonical tree type, e.g. "Pt<-1,0,0,Pt<0,0,-1>,Pt<1,0,-1>>" */
    int variant, boost, dbg, exact;
    const void* fn; /* host stub of the __global__ instantiation, for cudaLaunchKernel */
};

extern "C" void ps_aot_register(const PsAotEntry* table, int n);

#endif
