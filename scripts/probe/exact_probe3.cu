// probe3: device vs host exact stage on many near-degenerate inputs; build variants via -DB2_EXACT_HELPER=...
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../cgal_b200/csrc/predicates.cuh"
using namespace b2;
__global__ void k(const double* p, int n, int* o3, int* o2, int* oc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* q = p + 12 * i;
    o3[i] = exact_det3_sign(dd_diff(q[3], q[0]), dd_diff(q[4], q[1]), dd_diff(q[5], q[2]), dd_diff(q[6], q[0]), dd_diff(q[7], q[1]), dd_diff(q[8], q[2]),
                            dd_diff(q[9], q[0]), dd_diff(q[10], q[1]), dd_diff(q[11], q[2]));
    o2[i] = exact_det2_sign(dd_diff(q[3], q[0]), dd_diff(q[7], q[1]), dd_diff(q[6], q[0]), dd_diff(q[4], q[1]));
    oc[i] = exact_cmp_dist_sign(q, q + 3, q + 6);
}
static double rnd() { return rand() / (double)RAND_MAX; }
int main() {
    const int n = 200000;
    std::vector<double> rec(12 * n);
    srand(23);
    for (int i = 0; i < n; ++i) {
        double* q = &rec[12 * i];
        const double scale = (i % 3 == 0) ? 4000.0 : 1.0;
        for (int k = 0; k < 9; ++k) q[k] = scale * rnd();
        // 4th point: nearly coplanar combination of the first three (det ~ rounding level), or random
        const double u = rnd(), v = rnd();
        for (int k = 0; k < 3; ++k) q[9 + k] = q[k] + u * (q[3 + k] - q[k]) + v * (q[6 + k] - q[k]);
        if (i % 5 == 0) for (int k = 0; k < 3; ++k) q[9 + k] = scale * rnd();
        if (i % 7 == 0) { for (int k = 0; k < 3; ++k) q[6 + k] = q[3 + k]; q[6] += 1e-13 * scale; }   // near-equal distances for cmp_dist
    }
    double* d; int *o3, *o2, *oc;
    cudaMalloc(&d, rec.size() * 8); cudaMalloc(&o3, n * 4); cudaMalloc(&o2, n * 4); cudaMalloc(&oc, n * 4);
    cudaMemcpy(d, rec.data(), rec.size() * 8, cudaMemcpyHostToDevice);
    k<<<(n + 127) / 128, 128>>>(d, n, o3, o2, oc);
    std::vector<int> r3(n), r2(n), rc(n);
    cudaMemcpy(r3.data(), o3, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(r2.data(), o2, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(rc.data(), oc, n * 4, cudaMemcpyDeviceToHost);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    int bad3 = 0, bad2 = 0, badc = 0, zero3 = 0;
    for (int i = 0; i < n; ++i) {
        const double* q = &rec[12 * i];
        const int h3 = exact_det3_sign(dd_diff(q[3], q[0]), dd_diff(q[4], q[1]), dd_diff(q[5], q[2]), dd_diff(q[6], q[0]), dd_diff(q[7], q[1]), dd_diff(q[8], q[2]),
                                       dd_diff(q[9], q[0]), dd_diff(q[10], q[1]), dd_diff(q[11], q[2]));
        const int h2 = exact_det2_sign(dd_diff(q[3], q[0]), dd_diff(q[7], q[1]), dd_diff(q[6], q[0]), dd_diff(q[4], q[1]));
        const int hc = exact_cmp_dist_sign(q, q + 3, q + 6);
        bad3 += h3 != r3[i]; bad2 += h2 != r2[i]; badc += hc != rc[i]; zero3 += h3 == 0;
    }
    printf("n %d: det3 mismatches %d, det2 mismatches %d, cmp_dist mismatches %d (host det3 zeros %d)\n", n, bad3, bad2, badc, zero3);
    return bad3 + bad2 + badc ? 1 : 0;
}
