#include <cstdio>
#include "../../cgal_b200/csrc/predicates.cuh"
using namespace b2;
struct Trace { int n[8]; double v[8][192]; };
__host__ __device__ void trace_det3(const double* q, Trace& T) {
    DD ax = dd_diff(q[3], q[0]), ay = dd_diff(q[4], q[1]), az = dd_diff(q[5], q[2]), bx = dd_diff(q[6], q[0]), by = dd_diff(q[7], q[1]), bz = dd_diff(q[8], q[2]),
       cx = dd_diff(q[9], q[0]), cy = dd_diff(q[10], q[1]), cz = dd_diff(q[11], q[2]);
    T.n[0] = 18; double dd[18] = {ax.hi, ax.lo, ay.hi, ay.lo, az.hi, az.lo, bx.hi, bx.lo, by.hi, by.lo, bz.hi, bz.lo, cx.hi, cx.lo, cy.hi, cy.lo, cz.hi, cz.lo};
    for (int i = 0; i < 18; ++i) T.v[0][i] = dd[i];
    double m[16], t1[64], t2[64], t3[64], s[128], r[192];
    int nm = dd_det2(ax, by, ay, bx, m); T.n[1] = nm; for (int i = 0; i < nm; ++i) T.v[1][i] = m[i];
    const int n1 = exp_mul_dd(m, nm, cz, t1); T.n[2] = n1; for (int i = 0; i < n1; ++i) T.v[2][i] = t1[i];
    nm = dd_det2(ax, bz, az, bx, m); T.n[3] = nm; for (int i = 0; i < nm; ++i) T.v[3][i] = m[i];
    int n2 = exp_mul_dd(m, nm, cy, t2); for (int i = 0; i < n2; ++i) t2[i] = -t2[i]; T.n[4] = n2; for (int i = 0; i < n2; ++i) T.v[4][i] = t2[i];
    nm = dd_det2(ay, bz, az, by, m);
    const int n3 = exp_mul_dd(m, nm, cx, t3); T.n[5] = n3; for (int i = 0; i < n3; ++i) T.v[5][i] = t3[i];
    const int ns = exp_add(t1, n1, t2, n2, s); T.n[6] = ns; for (int i = 0; i < ns; ++i) T.v[6][i] = s[i];
    const int nr = exp_add(s, ns, t3, n3, r); T.n[7] = nr; for (int i = 0; i < nr; ++i) T.v[7][i] = r[i];
}
__global__ void k(const double* p, Trace* T) { trace_det3(p, *T); }
int main() {
    const double a[3] = {3948., 1806., 14.842434271702388}, b[3] = {3949., 1807., 14.743239361518222}, c[3] = {3948., 1807., 14.843862586086473};
    const double s[3] = {0., 0., 600.}, q[3] = {9.084787907804692e-01, 4.157081279688596e-01, 5.998653489007871e+02};
    double sh[3]; for (int i = 0; i < 3; ++i) sh[i] = a[i] + (q[i] - s[i]);
    double rec[12]; const double* quad[4] = {a, b, c, sh};
    for (int k2 = 0; k2 < 4; ++k2) for (int i = 0; i < 3; ++i) rec[3 * k2 + i] = quad[k2][i];
    double* d; Trace* dT; cudaMalloc(&d, sizeof(rec)); cudaMalloc(&dT, sizeof(Trace)); cudaMemset(dT, 0, sizeof(Trace));
    cudaMemcpy(d, rec, sizeof(rec), cudaMemcpyHostToDevice);
    k<<<1, 1>>>(d, dT);
    static Trace G, H; cudaMemcpy(&G, dT, sizeof(Trace), cudaMemcpyDeviceToHost);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    trace_det3(rec, H);
    const char* names[8] = {"dd", "m_xy", "t1", "m_xz", "t2", "t3", "s", "r"};
    for (int st = 0; st < 8; ++st) {
        printf("%s: n dev %d host %d\n", names[st], G.n[st], H.n[st]);
        int nn = G.n[st] > H.n[st] ? G.n[st] : H.n[st];
        for (int i = 0; i < nn; ++i) if (G.v[st][i] != H.v[st][i] || st < 2) printf("   [%d] dev %.17g host %.17g %s\n", i, G.v[st][i], H.v[st][i], G.v[st][i] != H.v[st][i] ? "<-- DIFF" : "");
    }
    return 0;
}
