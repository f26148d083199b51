// tables.h — host-side builders of the audio-independent K5 tables (per renderer, rebuilt on reconfigure):
//   need-list      texels of the smoothed R16 texture the module's shader can sample (lazy K5)
//   lazy tap table taps {index, weight} of those texels, tap-major for the L2 path and texel-major ("CSR" blob) for the
//                  shared-memory path of spectrum_kernel
//   full table     taps of ALL n texels per block of K5_BLOCK for k5_table_kernel
// Everything is evaluated with the code the kernels would run per tap (smooth_enumerate, gl_math.h: bit-identical on
// host and device).  No CUDA here: capi.cu uploads the vectors, tests/emul walks them on the host.
#ifndef GLB_TABLES_H
#define GLB_TABLES_H

#include "raster_core.h"

#include <cstring>
#include <vector>

namespace glb {

#ifndef GLB_TAPENTRY_DEFINED
#define GLB_TAPENTRY_DEFINED
struct alignas(8) TapEntry { int idx; float w; };   // one tap of the K5 smoothing sum: texel index, weight
#endif

// Lazy K5: the texels the module's fragment shader can sample, built with the SAME coordinate helpers the kernels
// use (raster_core.h).  Returns false when every texel may be needed (circle: continuous angle -> position).
inline bool build_need_list(const glava_b200_params& p, std::vector<int>* lists /* [2] */) {
    if (!p.smooth_pass || p.shader_pre_smoothed == 2) return false;   // (2: the shader smooths the K5 output again, any texel)
    std::vector<char> mark[2];
    mark[0].assign(p.n + 1, 0); mark[1].assign(p.n + 1, 0);
    auto hit = [&](int chan, float coord) {
        int i = (int) glm_rint(coord * (float) p.n);
        if (i >= 0 && i < p.n) mark[chan][i] = 1;
    };
    switch (p.module) {
        case GLAVA_B200_MOD_BARS: {
            int aw = p.bars_mirror_yx ? p.h : p.w;
            for (int x = 0; x < aw; ++x) {
                int chan; float pp; bool inner;
                if (bars_column_coord(p, (float) x + 0.5f, aw, &chan, &pp, &inner)) hit(chan, pp);
            }
            break;
        }
        case GLAVA_B200_MOD_RADIAL:
            for (int k = 0; k <= p.radial_nbars; ++k) {
                float pos = (float) k / (float) (p.radial_nbars / 2);
                hit(0, pos); hit(1, pos);
            }
            break;
        case GLAVA_B200_MOD_GRAPH: {
            float pixel = 1.0f / (float) p.w;
            for (int x = 0; x < p.w; ++x) {
                int chan; float c = graph_column_coord(p, x, &chan);
                hit(chan, g_max(c - pixel, 0.0f)); hit(chan, c); hit(chan, g_min(c + pixel, 1.0f));
            }
            if (p.graph_join_channels) {                 // `middle` (graph/1.frag:126): audio_l around 1, audio_r around 0
                hit(0, g_max(1.0f - pixel, 0.0f)); hit(0, 1.0f); hit(0, g_min(1.0f + pixel, 1.0f));
                hit(1, g_max(0.0f - pixel, 0.0f)); hit(1, 0.0f); hit(1, g_min(0.0f + pixel, 1.0f));
            }
            break;
        }
        case GLAVA_B200_MOD_WAVE:
            for (int x = -1; x <= p.w; ++x) mark[0][wave_tex_index(p.n, (float) x / (float) p.w)] = 1;
            break;
        case GLAVA_B200_MOD_TEST: break;            // samples nothing that reaches the output
        default: return false;
    }
    for (int c = 0; c < 2; ++c) {
        lists[c].clear();
        for (int i = 0; i < p.n; ++i) if (mark[c][i]) lists[c].push_back(i);
    }
    return true;
}

// ---- lazy K5 --------------------------------------------------------------------------------------------------------
struct LazyTables {
    size_t cnt = 0;                    // entries per channel (padded with -1 to the longer list)
    std::vector<int> need;             // [2][cnt] texel index or -1
    std::vector<int> tcnt;             // [2][cnt] taps of each entry
    std::vector<float> wsum;           // [2][cnt] sum of the weights in loop order
    size_t tap_max = 1;
    int epi_n = 0;                     // leading input bins any tap can reach
    std::vector<TapEntry> tab;         // [2][tap_max][cnt], tap-major (warp loads coalesce); unused slots {0, 0}
    // texel-major blob per channel: [float w[total]] [u16 idx[total]] [int off[cnt + 1]], sections 16-byte aligned
    std::vector<unsigned char> csr;    // [2][blob]
    size_t blob = 0, idx_off = 0, off_off = 0;
};

inline void build_lazy_tables(const glava_b200_params& p, const std::vector<int> lists[2], LazyTables* t) {
    size_t cnt = lists[0].size() > lists[1].size() ? lists[0].size() : lists[1].size();
    if (cnt == 0) cnt = 1;
    t->cnt = cnt;
    t->need.assign(2 * cnt, -1);
    for (int c = 0; c < 2; ++c) for (size_t i = 0; i < lists[c].size(); ++i) t->need[c * cnt + i] = lists[c][i];
    const SmoothParams sp = smooth_params(p);
    std::vector<std::vector<TapEntry>> taps(2 * cnt);
    t->wsum.assign(2 * cnt, 0.0f); t->tcnt.assign(2 * cnt, 0);
    t->tap_max = 1; t->epi_n = 0;
    for (size_t e = 0; e < 2 * cnt; ++e) {
        const int x = t->need[e];
        if (x < 0) continue;
        float weight = 0.0f;
        std::vector<TapEntry>& v = taps[e];
        smooth_enumerate(sp, p.n, ((float) x + 0.5f) / (float) p.n, [&](int i, float w) {
            weight += w;
            v.push_back(TapEntry { i, w });
        });
        t->wsum[e] = weight; t->tcnt[e] = (int) v.size();
        if (v.size() > t->tap_max) t->tap_max = v.size();
        for (const TapEntry& te : v) if (te.idx >= 0 && te.idx < p.n && te.idx + 1 > t->epi_n) t->epi_n = te.idx + 1;
    }
    // A tap outside the texture fetches 0 (texelFetch): texel * w = +0 either way, so it is stored as (index 0,
    // weight 0) and the kernels need no range test on the serial sum; its weight still counts in wsum.
    t->tab.assign(2 * t->tap_max * cnt, TapEntry { 0, 0.0f });
    for (size_t c = 0; c < 2; ++c)
        for (size_t k = 0; k < cnt; ++k) {
            const std::vector<TapEntry>& v = taps[c * cnt + k];
            for (size_t j = 0; j < v.size(); ++j) {
                const bool inside = v[j].idx >= 0 && v[j].idx < p.n;
                t->tab[(c * t->tap_max + j) * cnt + k] = inside ? v[j] : TapEntry { 0, 0.0f };
            }
        }
    size_t total = 0;
    for (size_t c = 0; c < 2; ++c) { size_t s = 0; for (size_t k = 0; k < cnt; ++k) s += taps[c * cnt + k].size(); if (s > total) total = s; }
    const size_t w_bytes = ((total * 4 + 15) / 16) * 16, i_bytes = ((total * 2 + 15) / 16) * 16, o_bytes = (((cnt + 1) * 4 + 15) / 16) * 16;
    t->blob = w_bytes + i_bytes + o_bytes; t->idx_off = w_bytes; t->off_off = w_bytes + i_bytes;
    t->csr.assign(2 * t->blob, 0);
    for (size_t c = 0; c < 2; ++c) {
        float* w = reinterpret_cast<float*>(t->csr.data() + c * t->blob);
        uint16_t* ix = reinterpret_cast<uint16_t*>(t->csr.data() + c * t->blob + t->idx_off);
        int* off = reinterpret_cast<int*>(t->csr.data() + c * t->blob + t->off_off);
        size_t at = 0;
        for (size_t k = 0; k < cnt; ++k) {
            off[k] = (int) at;
            for (const TapEntry& te : taps[c * cnt + k]) {
                const bool inside = te.idx >= 0 && te.idx < p.n;
                w[at] = inside ? te.w : 0.0f;
                ix[at] = inside ? (uint16_t) te.idx : (uint16_t) 0;
                ++at;
            }
        }
        off[cnt] = (int) at;
    }
}

// ---- need-list K5 out of shared memory (k5_need_smem_kernel) ----------------------------------------------------------
// Adjacent sampled texels have almost the same tap window (bars 1080p @4096: 159 texels per channel, 9.2 k taps, all inside
// 1536 bins): a block of texels whose windows' union fits a shared-memory tile is staged once per 32 streams and every
// texel of the block sums out of the tile — the L2 traffic of the tap loop drops by taps / union.
struct NeedBlocks {
    int nblk = 0;                  // blocks per channel (the shorter channel is padded with empty blocks)
    int max_rows = 0;              // tallest tile
    std::vector<int> blk;          // [2][nblk] x {first entry, end entry, first bin (even), rows (even)}
};
// target_rows <= 0: -target_rows per cent of the tallest single window (at least 192 rows)
inline void build_need_blocks(const LazyTables& t, int n, int target_rows, int texels_per_block, NeedBlocks* out) {
    std::vector<int> per[2];
    if (target_rows <= 0) {
        NeedBlocks single;
        build_need_blocks(t, n, 1, 1, &single);                  // one texel per block: max_rows = the tallest window
        const long long want = (long long) single.max_rows * (target_rows < 0 ? -target_rows : 105) / 100;
        target_rows = want > 192 ? (int) want : 192;
    }
    for (int c = 0; c < 2; ++c) {
        const float* w = reinterpret_cast<const float*>(t.csr.data() + c * t.blob);
        const uint16_t* ix = reinterpret_cast<const uint16_t*>(t.csr.data() + c * t.blob + t.idx_off);
        const int* off = reinterpret_cast<const int*>(t.csr.data() + c * t.blob + t.off_off);
        size_t k = 0;
        while (k < t.cnt && t.need[c * t.cnt + k] >= 0) {
            int lo = n, hi = -1, count = 0; const size_t k0 = k;
            while (k < t.cnt && t.need[c * t.cnt + k] >= 0 && count < texels_per_block) {
                int l = lo, h = hi;
                for (int o = off[k]; o < off[k + 1]; ++o) {
                    if (ix[o] == 0 && w[o] == 0.0f) continue;              // a tap outside the texture (or one that cannot count)
                    if ((int) ix[o] < l) l = ix[o];
                    if ((int) ix[o] > h) h = ix[o];
                }
                if (h < l) { l = l < n ? l : 0; h = h >= 0 ? h : 0; if (h < l) h = l; }
                const int rows = ((h - (l & ~1) + 1) + 1) & ~1;
                if (count > 0 && rows > target_rows) break;                  // (a single texel may exceed the target: its own block)
                lo = l; hi = h; ++count; ++k;
            }
            const int lo2 = lo & ~1, rows = ((hi - lo2 + 1) + 1) & ~1;
            per[c].push_back((int) k0); per[c].push_back((int) k); per[c].push_back(lo2); per[c].push_back(rows);
            if (rows > out->max_rows) out->max_rows = rows;
        }
    }
    const size_t nb = per[0].size() > per[1].size() ? per[0].size() / 4 : per[1].size() / 4;
    out->nblk = (int) nb;
    out->blk.assign(2 * nb * 4, 0);
    for (int c = 0; c < 2; ++c) for (size_t i = 0; i < per[c].size(); ++i) out->blk[c * nb * 4 + i] = per[c][i];
}

// ---- full-plane K5 ----------------------------------------------------------------------------------------------------
#ifndef K5_BLOCK
#define K5_BLOCK 128
#endif
struct alignas(16) K5Blk { int base, taps, lo, span; };   // first entry, taps per texel (padded), first input index, input span
struct alignas(8)  K5Ent { int idx; int wbits; };          // input index - lo, weight bits

struct K5TableHost {
    std::vector<K5Blk> blk;        // [ceil(n / K5_BLOCK)]
    std::vector<K5Ent> ent;        // per block [taps][K5_BLOCK], padded with {0, +0.0f}
    std::vector<float> wsum;       // [outputs]
    std::vector<int> out;          // [outputs] texel each output position writes
    int max_span = 1;
};

// Taps of the output texels `outs` (any subset of [0, n), in this order), in blocks of K5_BLOCK outputs.
// wsum / the kernel's thread index run over POSITIONS in `outs`; out[] maps a position back to its texel.
inline void build_k5_table_for(const glava_b200_params& p, const std::vector<int>& outs, K5TableHost* t) {
    const SmoothParams sp = smooth_params(p);
    const int n = p.n, cnt = (int) outs.size(), nblk = (cnt + K5_BLOCK - 1) / K5_BLOCK;
    t->blk.assign((size_t) (nblk > 0 ? nblk : 1), K5Blk { 0, 0, 0, 1 });
    t->ent.clear();
    t->wsum.assign((size_t) (cnt > 0 ? cnt : 1), 0.0f);
    t->out = outs;
    t->max_span = 1;
    std::vector<std::vector<TapEntry>> taps(K5_BLOCK);
    for (int b = 0; b < nblk; ++b) {
        int lo = n, hi = -1; size_t longest = 0;
        for (int k = 0; k < K5_BLOCK; ++k) {
            taps[k].clear();
            const int e = b * K5_BLOCK + k;
            if (e >= cnt) continue;
            const int x = outs[e];
            float weight = 0.0f;
            smooth_enumerate(sp, n, ((float) x + 0.5f) / (float) n, [&](int i, float w) {
                weight += w;
                // a tap outside the texture fetches 0: texel * w = +0 either way -> weight 0 (it still counts in wsum)
                const bool inside = i >= 0 && i < n;
                taps[k].push_back(TapEntry { inside ? i : -1, inside ? w : 0.0f });
                if (inside) { if (i < lo) lo = i; if (i > hi) hi = i; }
            });
            t->wsum[e] = weight;
            if (taps[k].size() > longest) longest = taps[k].size();
        }
        if (hi < lo) { lo = 0; hi = 0; }
        const int span = hi - lo + 1;
        if (span > t->max_span) t->max_span = span;
        const size_t base = t->ent.size();
        t->blk[b] = K5Blk { (int) base, (int) longest, lo, span };
        t->ent.resize(base + longest * K5_BLOCK, K5Ent { 0, 0 });              // padding: first staged texel, weight +0
        for (int k = 0; k < K5_BLOCK; ++k)
            for (size_t j = 0; j < taps[k].size(); ++j) {
                const TapEntry& te = taps[k][j];
                int wbits; memcpy(&wbits, &te.w, 4);
                t->ent[base + j * K5_BLOCK + k] = K5Ent { te.idx >= 0 ? te.idx - lo : 0, wbits };
            }
    }
}
inline void build_k5_table_host(const glava_b200_params& p, K5TableHost* t) {   // every texel of the plane
    std::vector<int> all((size_t) p.n);
    for (int i = 0; i < p.n; ++i) all[i] = i;
    build_k5_table_for(p, all, t);
}

}  // namespace glb

#endif
