// ORACLE — test infrastructure only.  Nothing here is linked into, imported by or executed from the
// product path (eao-fusion_b200/); only tests/, __graft_entry__.smoke() and bench.py's CPU legs call it.
//
// Independent, stage-by-stage CPU restatement of the reference ORB extractor on plain arrays.  Each
// function cites the reference lines it follows (/root/reference/...).  It is pinned by
// tests/test_oracle_vs_ref.py: byte-for-byte equality with oracle/_ref (the unmodified reference
// ORBextractor.cc compiled against cvshim) on synthetic and adversarial frames, and by the golden hashes
// in tests/golden/.  The reference ships no tests or golden vectors of its own, so parity
// is pinned on the reference's own code run here, not on reference fixtures.
//
// The quadtree (DistributeOctTree) is restated in the *parallel* formulation the CUDA kernel uses —
// keys never move, every key carries the list position of its node, passes relabel keys — so that the
// equality test against the reference's std::list implementation also validates the GPU algorithm.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../eao-fusion_b200/csrc/orb_pattern.h"
#include "cv_primitives.h"
#include "glibc_sincosf.h"

namespace {

const int kEdge = 19;       // EDGE_THRESHOLD   src/ORBextractor.cc:74
const int kHalfPatch = 15;  // HALF_PATCH_SIZE  src/ORBextractor.cc:73
const int kPatch = 31;      // PATCH_SIZE       src/ORBextractor.cc:72
const double kCvPi = 3.1415926535897932384626433832795;  // CV_PI

struct Tables {
    int nlevels;
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> quota;
    int umax[16];
};

// ORBextractor::ORBextractor  src/ORBextractor.cc:410-470
Tables make_tables(int nfeatures, float scaleFactorF, int nlevels) {
    Tables t;
    t.nlevels = nlevels;
    const double scaleFactor = scaleFactorF;  // member is double (include/ORBextractor.h:98)
    t.scale.assign(nlevels, 1.f);
    t.sigma2.assign(nlevels, 1.f);
    for (int i = 1; i < nlevels; ++i) {
        t.scale[i] = (float)(t.scale[i - 1] * scaleFactor);
        t.sigma2[i] = t.scale[i] * t.scale[i];
    }
    t.inv_scale.resize(nlevels);
    t.inv_sigma2.resize(nlevels);
    for (int i = 0; i < nlevels; ++i) {
        t.inv_scale[i] = 1.0f / t.scale[i];
        t.inv_sigma2[i] = 1.0f / t.sigma2[i];
    }
    t.quota.assign(nlevels, 0);
    const float factor = (float)(1.0f / scaleFactor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        t.quota[l] = cvprim::round_f(nDesired);
        sum += t.quota[l];
        nDesired *= factor;
    }
    t.quota[nlevels - 1] = std::max(nfeatures - sum, 0);
    // umax, :454-469
    const float hs = kHalfPatch * sqrtf(2.f) / 2;
    const int vmax = (int)floorf(hs + 1), vmin = (int)ceilf(hs);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) t.umax[v] = cvprim::round_d(sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (t.umax[v0] == t.umax[v0 + 1]) ++v0;
        t.umax[v] = v0;
        ++v0;
    }
    return t;
}

struct Level {
    int w, h;                  // inner size
    std::vector<uint8_t> buf;  // (w+38) x (h+38)
    int bw() const { return w + 2 * kEdge; }
    const uint8_t* inner() const { return buf.data() + (size_t)kEdge * bw() + kEdge; }
    uint8_t* inner() { return buf.data() + (size_t)kEdge * bw() + kEdge; }
};

// ORBextractor::ComputePyramid  src/ORBextractor.cc:1107-1132
void compute_pyramid(const uint8_t* img, int W, int H, size_t stride, const Tables& t, std::vector<Level>& pyr) {
    pyr.resize(t.nlevels);
    for (int l = 0; l < t.nlevels; ++l) {
        Level& L = pyr[l];
        L.w = cvprim::round_f((float)W * t.inv_scale[l]);
        L.h = cvprim::round_f((float)H * t.inv_scale[l]);
        L.buf.assign((size_t)L.bw() * (L.h + 2 * kEdge), 0);
        if (l == 0) {
            for (int y = 0; y < H; ++y) memcpy(L.inner() + (size_t)y * L.bw(), img + (size_t)y * stride, W);
        } else {
            cvprim::resize_linear_u8(pyr[l - 1].inner(), pyr[l - 1].w, pyr[l - 1].h, pyr[l - 1].bw(), L.inner(), L.w,
                                     L.h, L.bw());
        }
        cvprim::copy_make_border_reflect101(L.inner(), L.w, L.h, L.bw(), L.buf.data(), L.bw(), kEdge, kEdge, kEdge,
                                            kEdge);
    }
}

struct CellGrid {  // src/ORBextractor.cc:773-787
    int minB, maxBX, maxBY, nCols, nRows, wCell, hCell;
    bool valid;
};
CellGrid cell_grid(int cols, int rows) {
    CellGrid g;
    g.minB = kEdge - 3;
    g.maxBX = cols - kEdge + 3;
    g.maxBY = rows - kEdge + 3;
    const float width = (float)(g.maxBX - g.minB), height = (float)(g.maxBY - g.minB);
    g.nCols = (int)(width / 30.f);
    g.nRows = (int)(height / 30.f);
    g.valid = g.nCols > 0 && g.nRows > 0;
    g.wCell = g.valid ? (int)ceilf(width / g.nCols) : 0;
    g.hCell = g.valid ? (int)ceilf(height / g.nRows) : 0;
    return g;
}

struct Cand { int x, y, score; };  // window coordinates (origin = (16,16) of the level image)

// Per-cell FAST with ini/min retry  src/ORBextractor.cc:789-829 ; FAST semantics SURVEY.md A.4.
// Formulated the way the GPU does it: a threshold-independent arc score per pixel, then per-cell
// thresholding + 8-neighbour NMS that cannot see across the cell's inner area.
void fast_cells(const Level& L, int iniTh, int minTh, std::vector<Cand>& out) {
    out.clear();
    const CellGrid g = cell_grid(L.w, L.h);
    if (!g.valid) return;
    const int bw = L.bw();
    const uint8_t* im = L.inner();  // level coordinates
    std::vector<int> S;
    std::vector<Cand> cell;
    for (int i = 0; i < g.nRows; ++i) {
        const int iniY = g.minB + i * g.hCell;
        int maxY = iniY + g.hCell + 6;
        if (iniY >= g.maxBY - 3) continue;
        if (maxY > g.maxBY) maxY = g.maxBY;
        for (int j = 0; j < g.nCols; ++j) {
            const int iniX = g.minB + j * g.wCell;
            int maxX = iniX + g.wCell + 6;
            if (iniX >= g.maxBX - 6) continue;
            if (maxX > g.maxBX) maxX = g.maxBX;
            const int cw = maxX - iniX, ch = maxY - iniY;
            if (cw < 7 || ch < 7) continue;
            // arc score on the inner area [3,cw-3) x [3,ch-3)
            std::vector<int> B((size_t)cw * ch, -256);
            for (int y = 3; y < ch - 3; ++y)
                for (int x = 3; x < cw - 3; ++x)
                    B[(size_t)y * cw + x] = cvprim::fast_arc_best(im + (size_t)(iniY + y) * bw + iniX + x, bw);
            for (int pass = 0; pass < 2; ++pass) {
                const int th = pass == 0 ? iniTh : minTh;
                S.assign((size_t)cw * ch, 0);
                for (size_t k = 0; k < B.size(); ++k) S[k] = B[k] > th ? B[k] - 1 : 0;
                cell.clear();
                for (int y = 3; y < ch - 3; ++y)
                    for (int x = 3; x < cw - 3; ++x) {
                        if (!(B[(size_t)y * cw + x] > th)) continue;
                        const int* r = &S[(size_t)y * cw + x];
                        const int s = r[0];
                        if (s > r[-1] && s > r[1] && s > r[-cw - 1] && s > r[-cw] && s > r[-cw + 1] && s > r[cw - 1] &&
                            s > r[cw] && s > r[cw + 1])
                            cell.push_back({x + j * g.wCell, y + i * g.hCell, s});
                    }
                if (!cell.empty()) break;
            }
            out.insert(out.end(), cell.begin(), cell.end());
        }
    }
}

// ---------------------------------------------------------------------------------------------
// DistributeOctTree  src/ORBextractor.cc:539-763 (+ DivideNode :481-537), label formulation.
struct QNode {
    int x0, y0, x1, y1;  // UL.x, UL.y, UR.x, BR.y
    int cnt;
    bool noMore;
};

// Returns indices into `keys` in final list order.
void distribute_octree(const std::vector<Cand>& keys, int W, int H, int N, std::vector<int>& sel) {
    sel.clear();
    const int nk = (int)keys.size();
    const int nIni = (int)roundf((float)W / (float)H);
    if (nIni < 1) return;  // the reference indexes an empty vector here; callers reject such shapes
    const float hX = (float)W / nIni;

    std::vector<QNode> list;  // current list, front at index 0
    std::vector<int> label(nk);
    {
        std::vector<QNode> roots(nIni);
        for (int i = 0; i < nIni; ++i) roots[i] = {(int)(hX * (float)i), 0, (int)(hX * (float)(i + 1)), H, 0, false};
        std::vector<int> rootOf(nk);
        for (int k = 0; k < nk; ++k) {
            rootOf[k] = (int)((float)keys[k].x / hX);
            roots[rootOf[k]].cnt++;
        }
        std::vector<int> pos(nIni, -1);
        for (int i = 0; i < nIni; ++i)
            if (roots[i].cnt > 0) {  // :572-585
                roots[i].noMore = roots[i].cnt == 1;
                pos[i] = (int)list.size();
                list.push_back(roots[i]);
            }
        for (int k = 0; k < nk; ++k) label[k] = pos[rootOf[k]];
    }

    // One pass: split the nodes listed in `order` (list positions, in processing order), stopping after the
    // first split that brings the list size to >= stopAt (stopAt<0: never stop).  Children are created in
    // processing order n1..n4 and end up at the FRONT of the list in reverse creation order (push_front);
    // unsplit nodes keep their relative order behind them.  `cand` receives the new list positions of the
    // children with more than one key, in creation order.
    auto run_pass = [&](const std::vector<int>& order, int stopAt, std::vector<int>& cand) {
        const int n = (int)list.size();
        std::vector<int> cc((size_t)n * 4, 0);
        std::vector<char> wanted(n, 0);
        for (int p : order) wanted[p] = 1;
        for (int k = 0; k < nk; ++k) {
            const int p = label[k];
            if (!wanted[p]) continue;
            const QNode& nd = list[p];
            const int midX = nd.x0 + (int)ceilf((float)(nd.x1 - nd.x0) / 2);
            const int midY = nd.y0 + (int)ceilf((float)(nd.y1 - nd.y0) / 2);
            const int q = (keys[k].x < midX ? 0 : 1) + (keys[k].y < midY ? 0 : 2);
            cc[(size_t)p * 4 + q]++;
        }
        // sequential part: which prefix of `order` is actually split
        std::vector<char> split(n, 0);
        std::vector<int> firstChild(n, -1);
        int size = n, created = 0;
        for (int p : order) {
            int ne = 0;
            for (int q = 0; q < 4; ++q) ne += cc[(size_t)p * 4 + q] > 0;
            split[p] = 1;
            firstChild[p] = created;
            created += ne;
            size += ne - 1;
            if (stopAt >= 0 && size >= stopAt) break;
        }
        // new list
        std::vector<QNode> nl(size);
        std::vector<int> childPos((size_t)n * 4, -1), keepPos(n, -1);
        cand.clear();
        std::vector<std::pair<int, int>> candTmp;  // (creation index, new position)
        int kept = 0;
        for (int p = 0; p < n; ++p) {
            if (!split[p]) { keepPos[p] = created + kept; nl[created + kept] = list[p]; ++kept; }
        }
        for (int p = 0; p < n; ++p) {
            if (!split[p]) continue;
            const QNode& nd = list[p];
            const int hx = (int)ceilf((float)(nd.x1 - nd.x0) / 2), hy = (int)ceilf((float)(nd.y1 - nd.y0) / 2);
            const int mx = nd.x0 + hx, my = nd.y0 + hy;
            int c = firstChild[p];
            for (int q = 0; q < 4; ++q) {
                const int cnt = cc[(size_t)p * 4 + q];
                if (cnt == 0) continue;
                QNode ch;
                ch.x0 = (q & 1) ? mx : nd.x0;
                ch.x1 = (q & 1) ? nd.x1 : mx;
                ch.y0 = (q & 2) ? my : nd.y0;
                ch.y1 = (q & 2) ? nd.y1 : my;
                ch.cnt = cnt;
                ch.noMore = cnt == 1;
                const int np = created - 1 - c;  // push_front => reverse creation order
                nl[np] = ch;
                childPos[(size_t)p * 4 + q] = np;
                if (cnt > 1) candTmp.push_back({c, np});
                ++c;
            }
        }
        std::sort(candTmp.begin(), candTmp.end());
        for (auto& e : candTmp) cand.push_back(e.second);
        for (int k = 0; k < nk; ++k) {
            const int p = label[k];
            if (!split[p]) { label[k] = keepPos[p]; continue; }
            const QNode& nd = list[p];
            const int midX = nd.x0 + (int)ceilf((float)(nd.x1 - nd.x0) / 2);
            const int midY = nd.y0 + (int)ceilf((float)(nd.y1 - nd.y0) / 2);
            const int q = (keys[k].x < midX ? 0 : 1) + (keys[k].y < midY ? 0 : 2);
            label[k] = childPos[(size_t)p * 4 + q];
        }
        list.swap(nl);
    };

    bool finish = false;
    std::vector<int> cand, order;
    while (!finish) {
        const int prevSize = (int)list.size();
        order.clear();
        for (int p = 0; p < (int)list.size(); ++p)
            if (!list[p].noMore) order.push_back(p);  // sweep: front -> back, :606-665
        run_pass(order, -1, cand);
        const int nToExpand = (int)cand.size();
        if ((int)list.size() >= N || (int)list.size() == prevSize) {
            finish = true;
        } else if ((int)list.size() + nToExpand * 3 > N) {
            while (!finish) {  // :676-737
                const int prev2 = (int)list.size();
                // sort ascending by (size, creation index) and walk from the back.  `cand` is in creation
                // order and later-created nodes sit at SMALLER list positions, so the walk order is
                // (size desc, creation desc); SURVEY.md Appendix C-1 fixes address order == creation order.
                std::vector<std::pair<int, int>> s;  // (cnt, creation rank)
                for (int r = 0; r < (int)cand.size(); ++r) s.push_back({list[cand[r]].cnt, r});
                std::sort(s.begin(), s.end());
                order.clear();
                for (int j = (int)s.size() - 1; j >= 0; --j) order.push_back(cand[s[j].second]);
                std::vector<int> next;
                if (!order.empty()) run_pass(order, N, next);
                cand.swap(next);
                if ((int)list.size() >= N || (int)list.size() == prev2) finish = true;
            }
        }
    }
    // :741-760 — best response per node, first wins, keys scanned in input order
    std::vector<int> best(list.size(), -1);
    for (int k = 0; k < nk; ++k) {
        int& b = best[label[k]];
        if (b < 0 || keys[k].score > keys[b].score) b = k;
    }
    for (size_t p = 0; p < list.size(); ++p) sel.push_back(best[p]);
}

// IC_Angle  src/ORBextractor.cc:77-104 ; (x,y) in level coordinates on the bordered, unblurred level
float ic_angle(const Level& L, int x, int y, const int* umax) {
    const int step = L.bw();
    const uint8_t* c = L.inner() + (size_t)y * step + x;
    int m01 = 0, m10 = 0;
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int vsum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            const int p = c[u + v * step], m = c[u - v * step];
            vsum += p - m;
            m10 += u * (p + m);
        }
        m01 += v * vsum;
    }
    return cvprim::fast_atan2((float)m01, (float)m10);
}

// computeOrbDescriptor  src/ORBextractor.cc:107-147 ; img = blurred inner level (no border), step = w
void orb_descriptor(const uint8_t* img, int step, int x, int y, float angleDeg, uint8_t* desc) {
    const float factorPI = (float)(kCvPi / 180.f);  // :106
    const float angle = angleDeg * factorPI;
    float a, b;
    glibcf::sincosf_restated(angle, &b, &a);  // a = cosf(angle), b = sinf(angle)
    const uint8_t* c = img + (size_t)y * step + x;
    const signed char* p = kOrbPattern31;
    for (int i = 0; i < 32; ++i, p += 32) {
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            const float x0 = p[4 * k], y0 = p[4 * k + 1], x1 = p[4 * k + 2], y1 = p[4 * k + 3];
            const int r0 = cvprim::round_f(x0 * b + y0 * a), c0 = cvprim::round_f(x0 * a - y0 * b);
            const int r1 = cvprim::round_f(x1 * b + y1 * a), c1 = cvprim::round_f(x1 * a - y1 * b);
            const int t0 = c[r0 * step + c0], t1 = c[r1 * step + c1];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

}  // namespace

extern "C" {

struct eaoo_kp { float x, y, size, angle, response; int octave; };

// ---- primitives (pinned against cv2 in tests/test_oracle_primitives.py)
void eaoo_resize(const uint8_t* s, int sw, int sh, size_t ss, uint8_t* d, int dw, int dh, size_t ds) {
    cvprim::resize_linear_u8(s, sw, sh, ss, d, dw, dh, ds);
}
void eaoo_border(const uint8_t* s, int sw, int sh, size_t ss, uint8_t* d, size_t ds, int b) {
    cvprim::copy_make_border_reflect101(s, sw, sh, ss, d, ds, b, b, b, b);
}
int eaoo_fast(const uint8_t* img, int w, int h, size_t step, int th, int* xys, int cap) {
    std::vector<cvprim::FastPt> p;
    cvprim::fast9_nms(img, w, h, step, th, p);
    for (int i = 0; i < (int)p.size() && i < cap; ++i) { xys[3 * i] = p[i].x; xys[3 * i + 1] = p[i].y; xys[3 * i + 2] = p[i].score; }
    return (int)p.size();
}
void eaoo_blur(const uint8_t* s, int w, int h, size_t ss, uint8_t* d, size_t ds, int mode) {
    cvprim::gaussian_blur7(s, w, h, ss, d, ds, mode);
}
void eaoo_atan2(const float* y, const float* x, float* out, int n) {
    for (int i = 0; i < n; ++i) out[i] = cvprim::fast_atan2(y[i], x[i]);
}
void eaoo_sincosf(const float* x, float* s, float* c, long n) {
    for (long i = 0; i < n; ++i) glibcf::sincosf_restated(x[i], &s[i], &c[i]);
}
// compares the restatement with this host's libm over every `stride`-th float in [0, hi]; returns mismatches
long eaoo_sincosf_sweep(float hi, uint32_t stride) {
    uint32_t ul;
    memcpy(&ul, &hi, 4);
    long bad = 0;
    for (uint32_t u = 0; u <= ul; u += stride) {
        float f, s, c;
        memcpy(&f, &u, 4);
        glibcf::sincosf_restated(f, &s, &c);
        const float rs = sinf(f), rc = cosf(f);
        bad += memcmp(&s, &rs, 4) != 0;
        bad += memcmp(&c, &rc, 4) != 0;
    }
    return bad;
}

void eaoo_tables(int nfeatures, float sf, int nlevels, float* scale, float* inv, float* s2, float* is2, int* quota,
                 int* umax16) {
    const Tables t = make_tables(nfeatures, sf, nlevels);
    for (int i = 0; i < nlevels; ++i) {
        scale[i] = t.scale[i]; inv[i] = t.inv_scale[i]; s2[i] = t.sigma2[i]; is2[i] = t.inv_sigma2[i]; quota[i] = t.quota[i];
    }
    for (int i = 0; i < 16; ++i) umax16[i] = t.umax[i];
}

// ---- stages
// candidates of one bordered level (buffer (w+38)x(h+38)); returns count, writes (x,y,score) window coords
int eaoo_fast_cells(const uint8_t* bordered, int w, int h, int iniTh, int minTh, int* xys, int cap) {
    Level L;
    L.w = w; L.h = h;
    L.buf.assign(bordered, bordered + (size_t)(w + 38) * (h + 38));
    std::vector<Cand> c;
    fast_cells(L, iniTh, minTh, c);
    for (int i = 0; i < (int)c.size() && i < cap; ++i) { xys[3 * i] = c[i].x; xys[3 * i + 1] = c[i].y; xys[3 * i + 2] = c[i].score; }
    return (int)c.size();
}
// W,H = detection window size (maxBorder-minBorder); returns number selected; sel = indices in list order
int eaoo_octree(const int* xys, int n, int W, int H, int N, int* sel, int cap) {
    std::vector<Cand> k(n);
    for (int i = 0; i < n; ++i) k[i] = {xys[3 * i], xys[3 * i + 1], xys[3 * i + 2]};
    std::vector<int> s;
    distribute_octree(k, W, H, N, s);
    for (int i = 0; i < (int)s.size() && i < cap; ++i) sel[i] = s[i];
    return (int)s.size();
}

// Full ORBextractor::operator()  src/ORBextractor.cc:1043-1105.
// Optional dumps: pyr_out (bordered levels, concatenated), blur_out (inner blurred levels, concatenated),
// cand_out/cand_count (per-level candidates, level-major, cap cand_cap triples in total).
int eaoo_extract(const uint8_t* img, int W, int H, size_t stride, int nfeatures, float scaleFactor, int nlevels,
                 int iniTh, int minTh, int blur_mode, eaoo_kp* kps, uint8_t* desc, int cap, uint8_t* pyr_out,
                 uint8_t* blur_out, int* cand_out, int* cand_count, int cand_cap) {
    if (!img || W <= 0 || H <= 0) return -1;
    const Tables t = make_tables(nfeatures, scaleFactor, nlevels);
    std::vector<Level> pyr;
    compute_pyramid(img, W, H, stride, t, pyr);
    int n = 0, cand_total = 0;
    size_t pyr_off = 0, blur_off = 0;
    for (int l = 0; l < nlevels; ++l) {
        const Level& L = pyr[l];
        if (pyr_out) { memcpy(pyr_out + pyr_off, L.buf.data(), L.buf.size()); pyr_off += L.buf.size(); }
        std::vector<Cand> cands;
        fast_cells(L, iniTh, minTh, cands);
        if (cand_count) cand_count[l] = (int)cands.size();
        if (cand_out)
            for (const Cand& c : cands) {
                if (cand_total < cand_cap) { cand_out[3 * cand_total] = c.x; cand_out[3 * cand_total + 1] = c.y; cand_out[3 * cand_total + 2] = c.score; }
                ++cand_total;
            }
        const CellGrid g = cell_grid(L.w, L.h);
        std::vector<int> sel;
        distribute_octree(cands, g.maxBX - g.minB, g.maxBY - g.minB, t.quota[l], sel);
        std::vector<uint8_t> blur((size_t)L.w * L.h);
        const bool need_blur = !sel.empty() || blur_out;
        if (need_blur) cvprim::gaussian_blur7(L.inner(), L.w, L.h, L.bw(), blur.data(), L.w, blur_mode);
        if (blur_out) { memcpy(blur_out + blur_off, blur.data(), blur.size()); blur_off += blur.size(); }
        const int scaledPatch = (int)(kPatch * t.scale[l]);  // :835
        for (int s : sel) {
            const int x = cands[s].x + g.minB, y = cands[s].y + g.minB;  // :841-842
            const float ang = ic_angle(L, x, y, t.umax);
            if (n < cap) {
                eaoo_kp k;
                k.x = (float)x; k.y = (float)y;
                if (l != 0) { k.x *= t.scale[l]; k.y *= t.scale[l]; }  // :1095-1101
                k.size = (float)scaledPatch; k.angle = ang; k.response = (float)cands[s].score; k.octave = l;
                if (kps) kps[n] = k;
                if (desc) orb_descriptor(blur.data(), L.w, x, y, ang, desc + (size_t)n * 32);
            }
            ++n;
        }
    }
    return n;
}

}  // extern "C"
