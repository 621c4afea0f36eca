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
};

bool fast_fft_supported(int n);

// kind: 0 strided forward, 1 strided inverse, 2 strided forward + eigenvalue divide + inverse,
//       3 x real-to-complex (row pairs), 4 x complex-to-real
template <typename T>
int fast_pass(int kind, int N, T *R, void *W, const FastGeom &g, int tdim, const void *tw, const T *e0,
              const T *e1, const T *e2, T factor, int tile, Launch L);

}  // namespace insb
