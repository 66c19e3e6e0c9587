// TEST INFRASTRUCTURE: the pointwise DynamicSmagorinsky arithmetic of oceananigans.jl_b200/csrc/dynsmag.cuh compiled for the HOST
// (nvcc compiles `__host__ __device__` functions for both sides), so that tests/test_dynsmag_host.py can compare the exact code
// the kernels run with the numpy restatement of oracle/dynsmag.py on a machine without a GPU.  Float64 only.
#include "../oceananigans.jl_b200/csrc/dynsmag.cuh"

using namespace ob;

struct HostField {
    double *p;
    int Px, Py;
};
struct HostArgs {
    int N[3], H[3];
    double dx, dy, dz;
    const double *dzc, *rdzc, *rdzf;   // stretched z: arrays whose element [k + H[2] - 1 + 1]... see `koff`; nullptr when regular
    int koff;                          // element index of logical k = 0 in the three arrays
    HostField u, v, w, ub, vb, wb, Sg, Sb, LM, MM, nue;
    const double *J;                   // phase 4: [JLM (nout), JMM (nout)]
    int avg[3];
    double JLM_min;
};

static DField<double> view(const HostField &f, const int *H) {
    DField<double> d;
    d.p = f.p;
    d.sy = f.Px;
    d.sz = (long)f.Px * f.Py;
    d.off = (long)(H[0] - 1) + (long)(H[1] - 1) * d.sy + (long)(H[2] - 1) * d.sz;
    return d;
}

extern "C" int dynsmag_host(int phase, const HostArgs *a) {
    DynP<double> P;
    for (int d = 0; d < 3; d++) { P.N[d] = a->N[d]; P.H[d] = a->H[d]; }
    P.dx = a->dx; P.dy = a->dy; P.dz = a->dz;
    P.rdx = 1 / a->dx; P.rdy = 1 / a->dy; P.rdz = 1 / a->dz;
    P.dzc = a->dzc ? a->dzc + a->koff : nullptr;
    P.rdzc = a->rdzc ? a->rdzc + a->koff : nullptr;
    P.rdzf = a->rdzf ? a->rdzf + a->koff : nullptr;
    P.u = view(a->u, a->H); P.v = view(a->v, a->H); P.w = view(a->w, a->H);
    P.ub = view(a->ub, a->H); P.vb = view(a->vb, a->H); P.wb = view(a->wb, a->H);
    P.Sg = view(a->Sg, a->H); P.Sb = view(a->Sb, a->H); P.LM = view(a->LM, a->H); P.MM = view(a->MM, a->H);
    const DField<double> nue = view(a->nue, a->H);
    auto idx = [](const DField<double> &f, int i, int j, int k) { return f.off + i + (long)j * f.sy + (long)k * f.sz; };
    const int *N = a->N, *H = a->H;
    if (phase == 1) {
        for (int k = 2 - H[2]; k <= N[2] + H[2] - 1; k++)
            for (int j = 2 - H[1]; j <= N[1] + H[1] - 1; j++)
                for (int i = 2 - H[0]; i <= N[0] + H[0] - 1; i++) {
                    double x, y, z;
                    dyn_filter_velocities(P, i, j, k, x, y, z);
                    a->ub.p[idx(P.ub, i, j, k)] = x; a->vb.p[idx(P.vb, i, j, k)] = y; a->wb.p[idx(P.wb, i, j, k)] = z;
                }
        return 0;
    }
    for (int k = 1; k <= N[2]; k++)
        for (int j = 1; j <= N[1]; j++)
            for (int i = 1; i <= N[0]; i++) {
                if (phase == 2) {
                    double x, y;
                    dyn_sigma(P, i, j, k, x, y);
                    a->Sg.p[idx(P.Sg, i, j, k)] = x; a->Sb.p[idx(P.Sb, i, j, k)] = y;
                } else if (phase == 3) {
                    double x, y;
                    dyn_LM_MM(P, i, j, k, x, y);
                    a->LM.p[idx(P.LM, i, j, k)] = x; a->MM.p[idx(P.MM, i, j, k)] = y;
                } else {
                    const int nxo = a->avg[0] ? 1 : N[0], nyo = a->avg[1] ? 1 : N[1], nzo = a->avg[2] ? 1 : N[2];
                    const long nout = (long)nxo * nyo * nzo;
                    const long o = (a->avg[0] ? 0 : i - 1) + (long)nxo * ((a->avg[1] ? 0 : j - 1) + (long)nyo * (a->avg[2] ? 0 : k - 1));
                    a->nue.p[idx(nue, i, j, k)] = dyn_viscosity(P, i, j, k, a->J[o], a->J[nout + o], a->JLM_min);
                }
            }
    return 0;
}
