// K5: space transfer between nested FD grids (restriction / prolongation of mesh_to_mesh, TransferMesh.py:149-218).
// The reference multiplies with a sparse Kronecker product of 1-D interpolation matrices; here each 1-D operator is
// applied along its axis (kron(A, B) vec(G) = vec(A G B^T)), stored in ELL form (<= 8 entries per row: 6th-order
// interpolation has 6, full-weighting restriction 3).  One thread per pair of output points along the contiguous
// direction, all loads of a warp contiguous; the operators live in device memory (uploaded once per level pair).
#include "common.cuh"

namespace sdcb200 {
namespace {

struct AxisArgs {
    long long n_outer, n_inner;
    int n_out, width;
    const double* W;   // [n_out][width]
    const int* col;    // [n_out][width], -1 = unused slot
    const double* in;
    double* out;
    long long in_so, in_sa, out_so, out_sa;  // strides of the outer index and of the axis index (inner stride 1)
};

// out[o, i, c] = sum_t W[i, t] * in[o, col[i, t], c]
__global__ void __launch_bounds__(kThreads) axis_apply_kernel(const __grid_constant__ AxisArgs a) {
    const long long total = a.n_outer * a.n_out * a.n_inner;
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long e = (long long)blockIdx.x * kThreads + threadIdx.x; e < total; e += stride) {
        const long long c = e % a.n_inner;
        const long long oi = e / a.n_inner;
        const int i = (int)(oi % a.n_out);
        const long long o = oi / a.n_out;
        const double* src = a.in + o * a.in_so + c;
        double acc = 0.0;
        for (int t = 0; t < a.width; ++t) {
            const int k = a.col[i * a.width + t];
            if (k >= 0) acc = fma(a.W[i * a.width + t], src[(long long)k * a.in_sa], acc);
        }
        a.out[o * a.out_so + (long long)i * a.out_sa + c] = acc;
    }
}

}  // namespace
}  // namespace sdcb200

using namespace sdcb200;

extern "C" {

int sdcb200_axis_apply(long long n_outer, int n_out, long long n_inner, int width, const double* W_dev,
                       const int* col_dev, const double* in, long long in_stride_outer, long long in_stride_axis,
                       double* out, long long out_stride_outer, long long out_stride_axis, void* stream) {
    SDC_REQUIRE(n_outer >= 1 && n_out >= 1 && n_inner >= 1 && width >= 1 && width <= 16, "bad operator shape");
    SDC_REQUIRE(W_dev && col_dev && in && out, "null pointer");
    AxisArgs a;
    a.n_outer = n_outer;
    a.n_inner = n_inner;
    a.n_out = n_out;
    a.width = width;
    a.W = W_dev;
    a.col = col_dev;
    a.in = in;
    a.out = out;
    a.in_so = in_stride_outer;
    a.in_sa = in_stride_axis;
    a.out_so = out_stride_outer;
    a.out_sa = out_stride_axis;
    const long long total = n_outer * n_out * n_inner;
    long long blocks = (total + kThreads - 1) / kThreads;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    axis_apply_kernel<<<(int)blocks, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
    SDC_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
