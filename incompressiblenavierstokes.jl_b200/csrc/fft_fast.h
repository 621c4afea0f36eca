// fft_fast.h -- interface of the power-of-two FFT passes (fft_fast.cu).
#pragma once
#include "internal.h"

namespace insb {

struct FastGeom {
    int D;
    int n[3];         // Np
    int m0p;          // padded half-spectrum extent along x (multiple of the tile)
    long long ws[3];  // complex workspace strides
    long long rs[3];  // ghosted real strides
    // Strided passes only -- generalised addressing used by the slab-decomposed solve
    // (defaults reproduce the in-place pass on W with strides ws).  Element k of the line is at
    //   (k >> sh) * blk + (k & ((1 << sh) - 1)) * sd      (sh = 30: plain stride sd)
    // and the companion index `other` (blockIdx.y) contributes other * os.
    const void *in = nullptr;
    void *out = nullptr;
    long long in_sd = -1, in_os = -1, in_blk = 0;
    long long out_sd = -1, out_os = -1, out_blk = 0;
    int in_sh = 30, out_sh = 30;
    int nother = -1;   // extent of the companion index; -1 = n[other dim]
    int eig_off = 0;   // added to `other` for the eigenvalue lookup (ky offset of this rank)
    // Direct peer stores (slab-decomposed solve over NVLink P2P): when use_outp is set, block
    // (k >> out_sh) of the output goes to outp[k >> out_sh] (a peer-mapped buffer, already offset
    // to this rank's block there) instead of out + (k >> out_sh) * out_blk.
    // x R2C pass only: optional factor per row applied on load (the 1/Ω of psolver_tridiagonal,
    // pressure.jl:463-465), indexed by the row's index along dimension pre_dim (1 or 2)
    const void *pre = nullptr;
    int pre_dim = 1;
    // x R2C pass only: fused divergence! + scalewithvolume! (src/operators.jl:218-247).  When div_u
    // is set the real rows are computed on the fly from the velocity field (same evaluation order
    // as divergence_kernel, stencil.cu) instead of being read from R, which saves one write and
    // one read of the pressure array per solve.  div_u and the tables are offset to the first
    // pressure DOF like R; div_wrap: bit d = the lower neighbour of the first cell along the
    // periodic dimension d is the last cell (ghosts need not be filled).
    const void *div_u = nullptr;
    long long div_cs = 0;  // component stride of the velocity field
    const void *div_r[3] = {nullptr, nullptr, nullptr};  // guarded 1/Δ
    const void *div_d[3] = {nullptr, nullptr, nullptr};  // Δ
    int div_wrap = 0;
    // DCT passes (kinds 5-7): twiddles w_k and winv_k of manual_dct_stuff (src/utils.jl:65-87)
    const void *dw = nullptr, *dwi = nullptr;
    int use_outp = 0;
    // x R2C pass, decomposed psolver_tridiagonal: when x_kxl > 0 mode kx of row (y, z) is stored to
    // outp[kx / x_kxl] (the owner of that block of modes) at [z + x_zoff][y][kx % x_kxl]
    int x_kxl = 0, x_zoff = 0;
    // Row-grouped transpose layout (slab-decomposed solve): a [rows][m0p] plane is stored as
    // [rows/4][m0p/U][4][U] with U = 64 bytes of complex elements, so that the four consecutive
    // line elements a warp holds for one tile of U columns form ONE contiguous 256-byte run
    // (peer stores over NVLink reach 710 GB/s with >= 128-byte runs against 440-580 GB/s with
    // 64-byte ones, tools/p2p_microbench.cu).  0: plain strides, 1: the rows are the line
    // elements (within a block), 2: the rows are the companion index `other`.
    int in_grp = 0, out_grp = 0;
    void *outp[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // x C2R pass only: also store the periodic images of the result into the two ghost layers around
    // the interior (the internal layout of the fused time-step path, fused.h): bit 0 = x and y images,
    // bit 1 = z images (z periodic on this GPU)
    int img = 0;
};

bool fast_fft_supported(int n);

// kind: 0 strided forward, 1 strided inverse, 2 strided forward + eigenvalue divide + inverse,
//       3 x real-to-complex (row pairs), 4 x complex-to-real,
//       5 strided DCT-II, 6 strided DCT-III, 7 strided DCT-II + eigenvalue divide + DCT-III
template <typename T>
int fast_pass(int kind, int N, T *R, void *W, const FastGeom &g, int tdim, const void *tw, const T *e0,
              const T *e1, const T *e2, T factor, int tile, Launch L);

}  // namespace insb
