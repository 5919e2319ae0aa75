// host emulation of the patch addressing + packed basis: sum over the K-step schedule == direct dot product
#include "../rust-birdnet-onnx_b200/csrc/frontend_v24.h"
#include <cstdio>
#include <cstdlib>
#include <cmath>
using namespace bn;
static float h2f(uint16_t u) { __half_raw r; r.x = u; return __half2float(__half(r)); }
int main() {
    const int cfg[2][2] = {{2048, 278}, {1024, 280}};
    const int n_mels = 96, S = 144000, T = 511;
    SpecBranchHost hb[2];
    for (int b = 0; b < 2; ++b) if (!spec_v24_plan(cfg[b][0], cfg[b][1], n_mels, hb[b])) { printf("plan fail\n"); return 1; }
    int rp, ns; uint32_t plane, smem;
    if (!spec_v24_layout(hb, 2, rp, plane, ns, smem)) { printf("layout fail\n"); return 1; }
    printf("row_pitch %d plane %u stages %d smem %u ksteps %zu %zu\n", rp, plane, ns, smem, hb[0].table.size(), hb[1].table.size());
    std::vector<float> x(S);
    for (int i = 0; i < S; ++i) x[i] = (float)((i * 2654435761u >> 8) & 0xffff) / 65536.f - 0.5f;
    for (int b = 0; b < 2; ++b) {
        const int F = cfg[b][0], H = cfg[b][1], ldb = 96, N = hb[b].n_pad;
        std::vector<float> basis((size_t)F * ldb);
        for (size_t i = 0; i < basis.size(); ++i) basis[i] = (float)((i * 40503u >> 4) & 0xfff) / 4096.f - 0.5f;
        std::vector<uint16_t> pack; std::vector<uint32_t> tab;
        spec_v24_pack(hb[b], basis.data(), ldb, F, n_mels, pack);
        {   // the device's own enumeration: column pair outer, blocks inner (sv_group_steps)
            const int groups = (hb[b].kcells + 1) / 2;
            for (int m = 0; m < groups; ++m) {
                const int jn = (m < hb[b].split ? hb[b].blocks : hb[b].blocks - 1) + (m == groups - 1 ? hb[b].pad : 0);
                for (int j = 0; j < jn; ++j) tab.push_back(((uint32_t)(2 * m) * rp + j) * 16u);
            }
            if (tab.size() != hb[b].table.size()) { printf("schedule length mismatch %zu %zu\n", tab.size(), hb[b].table.size()); return 1; }
        }
        double worst = 0;
        for (int t0 : {0, 128, 384}) {
            // patch[c][r][8] floats
            std::vector<float> patch((size_t)36 * rp * 8, 0.f);
            for (int c = 0; c < hb[b].kcells; ++c)
                for (int r = 0; r < hb[b].rows; ++r)
                    for (int i = 0; i < 8; ++i) { int s = (t0 + r) * H + c * 8 + i; patch[((size_t)c * rp + r) * 8 + i] = s < S ? x[s] : 0.f; }
            for (int i : {0, 1, 77, 126, 127}) {
                const int t = t0 + i; if (t >= T) continue;
                for (int mel : {0, 5, 95}) {
                    double acc = 0;
                    for (size_t ks = 0; ks < tab.size(); ++ks) {
                        const uint32_t off = tab[ks] & 0xFFFFF;            // byte offset of (cell column, row j) : 16-byte cells
                        const uint32_t cell = off / 16;                    // = c * rp + j
                        for (int kc = 0; kc < 2; ++kc)
                            for (int kk = 0; kk < 8; ++kk) {
                                const float a = patch[((size_t)cell + (size_t)kc * rp + i) * 8 + kk];   // LBO = rp cells, row i = +i cells
                                const uint16_t* w = pack.data() + ks * (size_t)(2 * 2 * N * 8);
                                const float wv = h2f(w[((size_t)kc * 2 * N + mel) * 8 + kk]) + h2f(w[((size_t)kc * 2 * N + N + mel) * 8 + kk]);
                                acc += (double)a * wv;
                            }
                    }
                    double ref = 0;
                    for (int n = 0; n < F; ++n) ref += (double)x[t * H + n] * basis[(size_t)n * ldb + mel];
                    worst = fmax(worst, fabs(acc - ref));
                }
            }
        }
        printf("branch %d worst |schedule - direct| = %.3e\n", b, worst);
    }
    return 0;
}
