// probe: device exact stage on known inputs (expected signs from python fractions)
#include <cstdio>
#include "../../cgal_b200/csrc/predicates.cuh"
using namespace b2;
__global__ void k(const double* p, int n, int* out, double* dbg) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* q = p + 12 * i;
    out[i] = exact_det3_sign(dd_diff(q[3], q[0]), dd_diff(q[4], q[1]), dd_diff(q[5], q[2]), dd_diff(q[6], q[0]), dd_diff(q[7], q[1]), dd_diff(q[8], q[2]),
                             dd_diff(q[9], q[0]), dd_diff(q[10], q[1]), dd_diff(q[11], q[2]));
    if (i == 0) {
        DD ax = dd_diff(q[3], q[0]), by = dd_diff(q[7], q[1]), ay = dd_diff(q[4], q[1]), bx = dd_diff(q[6], q[0]);
        double m[16]; int nm = dd_det2(ax, by, ay, bx, m);
        dbg[0] = nm; for (int j = 0; j < nm; ++j) dbg[1 + j] = m[j];
        double h[8]; int nh = dd_mul(ax, by, h); dbg[20] = nh; for (int j = 0; j < nh; ++j) dbg[21 + j] = h[j];
        double x, y; two_prod(ax.hi, by.hi, x, y); dbg[30] = x; dbg[31] = y;
        two_sum(1e16, 1.0, x, y); dbg[32] = x; dbg[33] = y;
    }
}
int main() {
    const double a[3] = {3948., 1806., 14.842434271702388}, b[3] = {3949., 1807., 14.743239361518222}, c[3] = {3948., 1807., 14.843862586086473};
    const double s[3] = {0., 0., 600.}, q[3] = {9.084787907804692e-01, 4.157081279688596e-01, 5.998653489007871e+02};
    double sh[3]; for (int i = 0; i < 3; ++i) sh[i] = a[i] + (q[i] - s[i]);
    double rec[5][12]; const double* quads[5][4] = {{a, b, c, sh}, {a, b, c, s}, {s, q, a, b}, {s, q, b, c}, {s, q, c, a}};
    const int expect[5] = {-1, 1, -1, -1, 1};
    for (int r = 0; r < 5; ++r) for (int k = 0; k < 4; ++k) for (int i = 0; i < 3; ++i) rec[r][3 * k + i] = quads[r][k][i];
    double* d; int* o; double* dbg; cudaMalloc(&d, sizeof(rec)); cudaMalloc(&o, 20); cudaMalloc(&dbg, 64 * 8); cudaMemset(dbg, 0, 64 * 8);
    cudaMemcpy(d, rec, sizeof(rec), cudaMemcpyHostToDevice);
    k<<<1, 32>>>(d, 5, o, dbg);
    int res[5]; double hd[64]; cudaMemcpy(res, o, 20, cudaMemcpyDeviceToHost); cudaMemcpy(hd, dbg, 64 * 8, cudaMemcpyDeviceToHost);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    for (int r = 0; r < 5; ++r) printf("case %d: device %d expected %d %s\n", r, res[r], expect[r], res[r] == expect[r] ? "" : "  <-- WRONG");
    printf("det2 n=%g:", hd[0]); for (int j = 0; j < (int)hd[0]; ++j) printf(" %.17g", hd[1 + j]); printf("\n");
    printf("mul n=%g:", hd[20]); for (int j = 0; j < (int)hd[20]; ++j) printf(" %.17g", hd[21 + j]); printf("\n");
    printf("two_prod %.17g %.17g  two_sum %.17g %.17g\n", hd[30], hd[31], hd[32], hd[33]);
    // host evaluation of the same header
    for (int r = 0; r < 5; ++r) { const double* z = rec[r];
        printf("host %d\n", exact_det3_sign(dd_diff(z[3], z[0]), dd_diff(z[4], z[1]), dd_diff(z[5], z[2]), dd_diff(z[6], z[0]), dd_diff(z[7], z[1]), dd_diff(z[8], z[2]), dd_diff(z[9], z[0]), dd_diff(z[10], z[1]), dd_diff(z[11], z[2]))); }
    return 0;
}
