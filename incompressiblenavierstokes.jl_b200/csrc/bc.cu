// bc.cu -- ghost-cell boundary fills.
//
// Replaces apply_bc_u! / apply_bc_p! / apply_bc_temp! (src/boundary_conditions.jl:163-171,
// 200-208, 237-245) and their per-type methods (:273-285, :303-315, :341-372, :385, :388-402,
// :411-425, :442-450, :469-479, :494-499, :509-510).  Like the reference, dimensions are
// processed in order and each plane spans the FULL extent of the other dimensions, so edges
// and corners end up with exactly the reference's values.  One launch per dimension handles
// both sides and all components; the host turns the BC types into per-(side, component)
// "plane operations".
#include "internal.h"

namespace insb {

enum { OP_NONE = 0, OP_COPY = 1, OP_CONST = 2, OP_ARRAY = 3 };

struct PlaneOp {
    int mode, dst, src;
    double val;
    const void *arr;
};
constexpr int MAXBCF = 9;  // components / fields handled by one launch (9: a tensor field)
struct BcOps {
    PlaneOp op[2][MAXBCF];
    int ncomp;
};

template <typename T>
__global__ void __launch_bounds__(256)
    bc_plane_kernel(int D, int N0, int N1, int N2, long long cs, int beta, BcOps ops, T *f) {
    int side = blockIdx.y, comp = blockIdx.z;
    PlaneOp op = ops.op[side][comp];
    if (op.mode == OP_NONE) return;
    int N[3] = {N0, N1, N2};
    long long s[3] = {1, N0, (long long)N0 * N1};
    int d1 = beta == 0 ? 1 : 0;
    int d2 = beta == 2 ? 1 : 2;
    long long P = (long long)N[d1] * (D == 3 ? N[d2] : 1);
    long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q >= P) return;
    int i1 = (int)(q % N[d1]);
    int i2 = (int)(q / N[d1]);
    long long base = i1 * s[d1] + (D == 3 ? i2 * s[d2] : 0) + comp * cs;
    T v;
    if (op.mode == OP_COPY)
        v = f[base + op.src * s[beta]];
    else if (op.mode == OP_CONST)
        v = (T)op.val;
    else
        v = ((const T *)op.arr)[q];
    f[base + op.dst * s[beta]] = v;
}

template <typename T>
static int run_dim(const Grid<T> &g, int beta, const BcOps &ops, T *f, Launch L) {
    bool any = false;
    for (int s = 0; s < 2; s++)
        for (int c = 0; c < ops.ncomp; c++) any = any || ops.op[s][c].mode != OP_NONE;
    if (!any) return 0;
    int d1 = beta == 0 ? 1 : 0;
    int d2 = beta == 2 ? 1 : 2;
    long long P = (long long)g.N[d1] * (g.D == 3 ? g.N[d2] : 1);
    dim3 gr((unsigned)((P + 255) / 256), 2, ops.ncomp);
    bc_plane_kernel<T><<<gr, 256, 0, L.stream>>>(g.D, g.N[0], g.N[1], g.N[2], g.cs, beta, ops, f);
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}

// DOF offsets (src/boundary_conditions.jl:83-93)
static int offset_u(int bc, bool isright, bool isnormal) {
    switch (bc) {
        case INSB200_BC_PERIODIC: return 1;
        case INSB200_BC_DIRICHLET:
        case INSB200_BC_SYMMETRIC: return 1 + (isright && isnormal ? 1 : 0);
        default: return 1 + (!isright && !isnormal ? 1 : 0);
    }
}
static int offset_p(int bc, bool isright) {
    return bc == INSB200_BC_PRESSURE ? 1 + (!isright ? 1 : 0) : 1;
}

template <typename T>
static int halo_dim(const Grid<T> &g, const BcDesc &b, T *f, int ncomp, Launch L) {
    if (L.counter) *L.counter += comm_halo_launches(b.comm);
    const bool ring = b.bc_u[2][0] == INSB200_BC_PERIODIC;  // walls / symmetry planes in z: open ends
    return comm_halo_z(b.comm, f, sizeof(T), g.s[2], g.N[2], ncomp, g.cs, L.stream, ring);
}

// Slab decomposition along a NON-periodic z (walls / symmetry planes, BASELINE config 4): the interfaces between
// slabs are exchanged like periodic ghosts; the physical boundary planes exist on the first rank (lower side) and
// on the last rank (upper side) only, where the plane operations of the reference run on the local slab (the
// exchange is open-ended: nothing travels between the last and the first rank).
static bool z_decomposed_walls(const BcDesc &b, const int (*types)[2]) {
    return b.comm && (types[2][0] != INSB200_BC_PERIODIC || types[2][1] != INSB200_BC_PERIODIC);
}
static void keep_end_rank_sides(const BcDesc &b, BcOps &ops) {
    const int rank = comm_rank(b.comm), world = comm_world(b.comm);
    for (int c = 0; c < MAXBCF; c++) {
        if (rank != 0) ops.op[0][c].mode = OP_NONE;
        if (rank != world - 1) ops.op[1][c].mode = OP_NONE;
    }
}

template <typename T>
int launch_bc_u(const Grid<T> &g, const BcDesc &b, T *u, int dudt, Launch L, int dims) {
    for (int beta = 0; beta < g.D; beta++) {
        if (!(dims >> beta & 1)) continue;
        const bool slab_dim = beta == 2 && b.comm;
        if (slab_dim) {
            INSB_TRY(halo_dim<T>(g, b, u, g.D, L));
            if (!z_decomposed_walls(b, b.bc_u)) continue;
        }
        BcOps ops{};
        ops.ncomp = g.D;
        int n = g.N[beta];
        int plo = offset_p(b.bc_u[beta][0], false), phi = n - offset_p(b.bc_u[beta][1], true);
        for (int side = 0; side < 2; side++) {
            int bc = b.bc_u[beta][side];
            bool isright = side == 1;
            for (int c = 0; c < g.D; c++) {
                PlaneOp &op = ops.op[side][c];
                bool normal = c == beta;
                int ulo = offset_u(b.bc_u[beta][0], false, normal);
                int uhi = n - offset_u(b.bc_u[beta][1], true, normal);
                int plane = isright ? uhi : ulo - 1;  // boundary(β, N, Iu[α], isright)
                switch (bc) {
                    case INSB200_BC_PERIODIC:  // :273-285 (planes from Ip)
                        op.mode = OP_COPY;
                        op.dst = isright ? phi : plo - 1;
                        op.src = isright ? plo : phi - 1;
                        break;
                    case INSB200_BC_DIRICHLET:  // :341-372
                        if (b.u_plane[beta][side][c]) {
                            op.mode = OP_ARRAY;
                            op.arr = b.u_plane[beta][side][c];
                        } else {
                            op.mode = OP_CONST;
                            op.val = dudt ? 0.0 : b.u_val[beta][side][c];
                        }
                        op.dst = plane;
                        break;
                    case INSB200_BC_SYMMETRIC:  // :411-425
                        op.dst = plane;
                        if (normal) {
                            op.mode = OP_CONST;
                            op.val = 0.0;
                        } else {
                            op.mode = OP_COPY;
                            op.src = isright ? plane - 1 : plane + 1;
                        }
                        break;
                    case INSB200_BC_PRESSURE:  // :469-479
                        op.mode = OP_COPY;
                        op.dst = plane;
                        op.src = isright ? plane - 1 : plane + 1;
                        break;
                }
            }
        }
        if (slab_dim) keep_end_rank_sides(b, ops);
        INSB_TRY(run_dim<T>(g, beta, ops, u, L));
    }
    return 0;
}

// nf scalar fields of the same kind back to back (stride g.cs) share the launches
template <typename T>
static int scalar_bc(const Grid<T> &g, const BcDesc &b, const int (*types)[2], bool is_temp, T *p,
                     Launch L, int dims = 7, int nf = 1) {
    for (int beta = 0; beta < g.D; beta++) {
        if (!(dims >> beta & 1)) continue;
        const bool slab_dim = beta == 2 && b.comm;
        if (slab_dim) {
            for (int f0 = 0; f0 < nf; f0 += 3)  // (three planes per side and exchange)
                INSB_TRY(halo_dim<T>(g, b, p + (long long)f0 * g.cs, nf - f0 < 3 ? nf - f0 : 3, L));
            if (!z_decomposed_walls(b, b.bc_u) && !z_decomposed_walls(b, types)) continue;
        }
        BcOps ops{};
        ops.ncomp = nf;
        int n = g.N[beta];
        // planes always come from Ip, which is defined by the VELOCITY BC types
        int plo = offset_p(b.bc_u[beta][0], false), phi = n - offset_p(b.bc_u[beta][1], true);
        for (int side = 0; side < 2; side++) {
            int bc = types[beta][side];
            bool isright = side == 1;
            PlaneOp &op = ops.op[side][0];
            int plane = isright ? phi : plo - 1;
            switch (bc) {
                case INSB200_BC_PERIODIC:  // :303-315
                    op.mode = OP_COPY;
                    op.dst = plane;
                    op.src = isright ? plo : phi - 1;
                    break;
                case INSB200_BC_DIRICHLET:  // p: no-op (:385); temp: constant (:388-402)
                    if (is_temp) {
                        op.mode = OP_CONST;
                        op.dst = plane;
                        op.val = b.temp_val[beta][side];
                    }
                    break;
                case INSB200_BC_SYMMETRIC:  // :442-450
                    op.mode = OP_COPY;
                    op.dst = plane;
                    op.src = isright ? plane - 1 : plane + 1;
                    break;
                case INSB200_BC_PRESSURE:  // p: zero (:494-499); temp: symmetric (:509-510)
                    op.dst = plane;
                    if (is_temp) {
                        op.mode = OP_COPY;
                        op.src = isright ? plane - 1 : plane + 1;
                    } else {
                        op.mode = OP_CONST;
                        op.val = 0.0;
                    }
                    break;
            }
        }
        for (int side = 0; side < 2; side++)
            for (int f = 1; f < nf; f++) ops.op[side][f] = ops.op[side][0];
        if (slab_dim) keep_end_rank_sides(b, ops);
        INSB_TRY(run_dim<T>(g, beta, ops, p, L));
    }
    return 0;
}

// apply_bc_p! on nf scalar fields stored back to back (the tensor components of les.cu)
template <typename T>
int launch_bc_p_fields(const Grid<T> &g, const BcDesc &b, T *fields, int nf, Launch L) {
    return scalar_bc<T>(g, b, b.bc_u, false, fields, L, 7, nf);
}

// zero_out_wall! (src/eddyviscosity.jl:344-353): first and last plane of every Dirichlet dimension
template <typename T>
int launch_zero_out_wall(const Grid<T> &g, const BcDesc &b, T *fields, int nf, Launch L) {
    for (int beta = 0; beta < g.D; beta++) {
        const bool slab_dim = beta == 2 && b.comm;
        if (slab_dim && !z_decomposed_walls(b, b.bc_u)) continue;  // a periodic decomposed direction has no wall
        BcOps ops{};
        ops.ncomp = nf;
        for (int side = 0; side < 2; side++)
            if (b.bc_u[beta][side] == INSB200_BC_DIRICHLET)
                for (int f = 0; f < nf; f++) {
                    ops.op[side][f].mode = OP_CONST;
                    ops.op[side][f].val = 0.0;
                    ops.op[side][f].dst = side ? g.N[beta] - 1 : 0;
                }
        if (slab_dim) keep_end_rank_sides(b, ops);
        INSB_TRY(run_dim<T>(g, beta, ops, fields, L));
    }
    return 0;
}

template <typename T>
int launch_bc_p(const Grid<T> &g, const BcDesc &b, T *p, Launch L, int dims) {
    return scalar_bc<T>(g, b, b.bc_u, false, p, L, dims);
}
template <typename T>
int launch_bc_temp(const Grid<T> &g, const BcDesc &b, T *temp, Launch L) {
    return scalar_bc<T>(g, b, b.bc_temp, true, temp, L);
}

template int launch_bc_u<float>(const Grid<float> &, const BcDesc &, float *, int, Launch, int);
template int launch_bc_u<double>(const Grid<double> &, const BcDesc &, double *, int, Launch, int);
template int launch_bc_p<float>(const Grid<float> &, const BcDesc &, float *, Launch, int);
template int launch_bc_p<double>(const Grid<double> &, const BcDesc &, double *, Launch, int);
template int launch_bc_p_fields<float>(const Grid<float> &, const BcDesc &, float *, int, Launch);
template int launch_bc_p_fields<double>(const Grid<double> &, const BcDesc &, double *, int, Launch);
template int launch_zero_out_wall<float>(const Grid<float> &, const BcDesc &, float *, int, Launch);
template int launch_zero_out_wall<double>(const Grid<double> &, const BcDesc &, double *, int, Launch);
template int launch_bc_temp<float>(const Grid<float> &, const BcDesc &, float *, Launch);
template int launch_bc_temp<double>(const Grid<double> &, const BcDesc &, double *, Launch);

}  // namespace insb
