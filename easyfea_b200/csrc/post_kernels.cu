// Post-processing fields: Hooke's law at the Gauss points, per-element strain / stress results
// (components, von Mises, whole field), deformation energy per element and element -> node averaging.  All of them stream
// per-Gauss-point fields once (HBM-bound, one thread per element / node); the strain field itself comes from efb_strain.
#include "common.cuh"

namespace efb {

// sigma = C eps at every Gauss point        Models/Elastic/_laws.py:159-185 (C homogeneous, per element or per Gauss point)
__global__ void k_hooke(long long n_gp, int nPg, int ns, const double* __restrict__ eps, const double* __restrict__ C, int C_mode,
                        double* __restrict__ sigma) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_gp; i += (long long)gridDim.x * blockDim.x) {
        const double* Cp = C + (C_mode == 0 ? 0 : (C_mode == 1 ? (i / nPg) : i) * (long long)(ns * ns));
        const double* e = eps + i * ns;
        double ev[6];
        for (int k = 0; k < ns; ++k) ev[k] = e[k];
        for (int r = 0; r < ns; ++r) {
            double s = 0.0;
            for (int k = 0; k < ns; ++k) s += Cp[r * ns + k] * ev[k];
            sigma[i * ns + r] = s;
        }
    }
}

// `__Result_in_Strain_or_Stress_field(...).mean(1)`, Models/_utils.py:302-430: shear components are divided by `coef`, then
// what >= 0: that component; what == -1: von Mises; what == -2: the whole (rescaled) field.  out (Ne) or (Ne, ns).
__global__ void k_field_result(long long Ne, int nPg, int dim, const double* __restrict__ field, int what, double coef,
                               double* __restrict__ out) {
    const int ns = dim == 2 ? 3 : 6;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < Ne; e += (long long)gridDim.x * blockDim.x) {
        double acc[6] = {0, 0, 0, 0, 0, 0};
        for (int p = 0; p < nPg; ++p) {
            const double* f = field + (e * nPg + p) * ns;
            double v[6];
            for (int k = 0; k < ns; ++k) v[k] = (k < dim) ? f[k] : f[k] * (1.0 / coef);
            if (what == -2) {
                for (int k = 0; k < ns; ++k) acc[k] += v[k];
            } else if (what == -1) {
                double vm;
                if (dim == 2) {
                    const double xx = v[0], yy = v[1], xy = v[2];
                    vm = sqrt(xx * xx + yy * yy - xx * yy + 3.0 * (xy * xy));
                } else {
                    const double xx = v[0], yy = v[1], zz = v[2], yz = v[3], xz = v[4], xy = v[5];
                    vm = sqrt(0.5 * ((xx - yy) * (xx - yy) + (yy - zz) * (yy - zz) + (zz - xx) * (zz - xx) +
                                     6.0 * (xy * xy + yz * yz + xz * xz)));
                }
                acc[0] += vm;
            } else {
                acc[0] += v[what];
            }
        }
        if (what == -2) {
            for (int k = 0; k < ns; ++k) out[e * ns + k] = acc[k] / nPg;
        } else {
            out[e] = acc[0] / nPg;
        }
    }
}

// Wdef_e = scale * sum_p wJ * 1/2 sigma.eps          Simulations/_elastic.py:323-396 (`_Calc_Psi_Elas`)
__global__ void k_energy(long long Ne, int nPg, int ns, const double* __restrict__ eps, const double* __restrict__ sigma,
                         const double* __restrict__ wJ, double scale, double* __restrict__ out) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < Ne; e += (long long)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int p = 0; p < nPg; ++p) {
            const long long i = e * nPg + p;
            double s = 0.0;
            for (int k = 0; k < ns; ++k) s += sigma[i * ns + k] * eps[i * ns + k];
            acc += scale * wJ[i] * (0.5 * s);
        }
        out[e] = acc;
    }
}

// `Mesh.Get_Node_Values`, FEM/_mesh.py:822-873: node value = (sum of the values of the elements around the node) / their count,
// summed in ascending element order like scipy's `connect_n_e @ values_e`.  rowptr/qlist = the assembly pattern's node ->
// element-row lists (q = e * nPe + a, ascending), one element group.
__global__ void k_node_values(long long Nn, const long long* __restrict__ rowptr, const long long* __restrict__ qlist, int nPe,
                              const double* __restrict__ values_e, int ncols, double* __restrict__ out) {
    const long long total = Nn * ncols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / ncols;
        const int c = (int)(i - n * ncols);
        const long long r0 = rowptr[n], r1 = rowptr[n + 1];
        double s = 0.0;
        for (long long r = r0; r < r1; ++r) s += values_e[(qlist[r] / nPe) * ncols + c];
        out[i] = r1 > r0 ? s / (double)(r1 - r0) : 0.0;
    }
}

static unsigned grid_for(long long n) {
    long long b = (n + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace efb

using namespace efb;

extern "C" int efb_hooke(int64_t Ne, int nPg, int ns, const double* eps, const double* C, int C_mode, double* sigma, void* stream) {
    if ((ns != 3 && ns != 6) || C_mode < 0 || C_mode > 2) {
        set_error("efb_hooke: ns must be 3 or 6 and C_mode 0 (homogeneous), 1 (per element) or 2 (per Gauss point)");
        return 1;
    }
    if (Ne == 0) return 0;
    k_hooke<<<grid_for(Ne * nPg), 256, 0, as_stream(stream)>>>(Ne * (long long)nPg, nPg, ns, eps, C, C_mode, sigma);
    return check_launch("efb_hooke");
}

extern "C" int efb_field_result(int64_t Ne, int nPg, int dim, const double* field, int what, double coef, double* out, void* stream) {
    const int ns = dim == 2 ? 3 : 6;
    if ((dim != 2 && dim != 3) || what < -2 || what >= ns || coef == 0.0) {
        set_error("efb_field_result: dim 2 or 3, what in [-2, %d), coef != 0", ns);
        return 1;
    }
    if (Ne == 0) return 0;
    k_field_result<<<grid_for(Ne), 256, 0, as_stream(stream)>>>(Ne, nPg, dim, field, what, coef, out);
    return check_launch("efb_field_result");
}

extern "C" int efb_energy_e(int64_t Ne, int nPg, int ns, const double* eps, const double* sigma, const double* wJ, double scale,
                            double* out, void* stream) {
    if (Ne == 0) return 0;
    k_energy<<<grid_for(Ne), 256, 0, as_stream(stream)>>>(Ne, nPg, ns, eps, sigma, wJ, scale, out);
    return check_launch("efb_energy_e");
}

extern "C" int efb_node_values(int64_t Nn, const int64_t* rowptr, const int64_t* qlist, int nPe, const double* values_e, int ncols,
                               double* out, void* stream) {
    if (Nn == 0 || ncols == 0) return 0;
    k_node_values<<<grid_for(Nn * ncols), 256, 0, as_stream(stream)>>>(Nn, (const long long*)rowptr, (const long long*)qlist, nPe, values_e,
                                                                       ncols, out);
    return check_launch("efb_node_values");
}
