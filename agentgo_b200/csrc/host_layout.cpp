// host_layout.cpp — host side of the device layouts (kernel_args.h): start positions re-indexed for
// the warp kernel (PaddedPosition) and the group kernel (GroupPosition), and the GF(2) jump table of
// TinyMT32.  Plain C++ (no device code), so the test-only emulator harness links it too.
#include "kernel_args.h"

namespace agb {

constexpr int kBorder = 3;

// Host side of the padded layout: re-index a Position (board_ops.h) into a PaddedPosition.
void to_padded(const Position& p, PaddedPosition* out) {
    const int n = p.n, w = n + 2;
    auto pad = [&](int idx) { return (idx / n + 1) * w + (idx % n + 1); };
    for (int i = 0; i < kMaxNP; i++) { out->w0[i] = (uint32_t)kBorder << 30; out->w1[i] = 0; }
    for (int idx = 0; idx < n * n; idx++) {
        const uint32_t v0 = p.w0[idx], v1 = p.w1[idx];
        const int col = (int)((v0 >> 20) & 3u);
        const int a = (int)(v0 & 1023u), b = (int)((v0 >> 10) & 1023u);
        if (col == 0) { out->w0[pad(idx)] = 0; continue; }         // the slot is filled in below
        out->w0[pad(idx)] = (uint32_t)pad(a) | ((uint32_t)(b == kNil ? kNilP : pad(b)) << 18) | ((uint32_t)col << 30);
        const bool is_root = a == idx;                              // hp/tail are meaningful at roots only
        out->w1[pad(idx)] = is_root ? ((v1 & 0xffffu) | ((uint32_t)pad((int)(v1 >> 16)) << 16)) : 0u;
    }
    // the reference's empty list, head first (self-link = last, Board.cpp:46-54), becomes the
    // array elist read from its end
    int order[kMaxNN], m = 0;
    for (int idx = p.head; idx >= 0 && m < n * n;) {
        order[m++] = idx;
        const int nx = (int)((p.w0[idx] >> 10) & 1023u);
        if (nx == idx) break;
        idx = nx;
    }
    uint16_t* el = reinterpret_cast<uint16_t*>(out->elist_w);
    for (int i = 0; i < (int)(sizeof(out->elist_w) / 2); i++) el[i] = 0;
    for (int k = 0; k < m; k++) {
        const int c = pad(order[m - 1 - k]);
        el[k] = (uint16_t)c;
        out->w0[c] = (uint32_t)k;
    }
    out->n_el = (int16_t)m;
    out->stones = (int16_t)(p.num_black + p.num_white);
    out->ko_key_col = (int16_t)(p.ko ? p.ko_col : 0);
    out->ko_idx = (int16_t)(p.ko ? pad(p.ko_idx) : 0);
    out->last[0] = (int16_t)(p.last[0] >= 0 ? pad(p.last[0]) : -1);
    out->last[1] = (int16_t)(p.last[1] >= 0 ? pad(p.last[1]) : -1);
    out->ending = p.ending;
    out->pad = 0;
}

// The one-word layout of the group kernel (kernel_args.h).
void to_group(const Position& p, GroupPosition* out) {
    // one border column: rows of n + 1 points (GCfg in playout_group.cuh)
    const int n = p.n, w = n + 1;
    auto pad = [&](int idx) { return (idx / n + 1) * w + (idx % n + 1); };
    for (int i = 0; i < kMaxNP; i++) out->w[i] = gw::border((uint32_t)i);
    for (int idx = 0; idx < n * n; idx++) {
        const uint32_t v0 = p.w0[idx], v1 = p.w1[idx];
        const int col = (int)((v0 >> 20) & 3u);
        const int a = (int)(v0 & 1023u), b = (int)((v0 >> 10) & 1023u);
        if (col == 0) { out->w[pad(idx)] = gw::empty((uint32_t)pad(idx), 0); continue; }   // the slot is filled in below
        const bool is_root = a == idx;
        out->w[pad(idx)] = gw::stone((uint32_t)pad(a), (uint32_t)(b == kNil ? 511 : pad(b)), is_root ? (v1 & 0x7ffu) : 0u,
                                     (uint32_t)col);
    }
    // the second stone of every string names the string's last stone
    for (int idx = 0; idx < n * n; idx++) {
        const uint32_t v0 = p.w0[idx];
        if (((v0 >> 20) & 3u) == 0 || (int)(v0 & 1023u) != idx) continue;      // roots only
        const int second = (int)((v0 >> 10) & 1023u);
        if (second == kNil) continue;
        const int tail = (int)(p.w1[idx] >> 16);
        out->w[pad(second)] = (out->w[pad(second)] & ~(0x7ffu << gw::kFShift)) | ((uint32_t)pad(tail) << gw::kFShift);
    }
    int order[kMaxNN], m = 0;
    for (int idx = p.head; idx >= 0 && m < n * n;) {
        order[m++] = idx;
        const int nx = (int)((p.w0[idx] >> 10) & 1023u);
        if (nx == idx) break;
        idx = nx;
    }
    uint16_t* el = reinterpret_cast<uint16_t*>(out->elist_w);
    for (int i = 0; i < (int)(sizeof(out->elist_w) / 2); i++) el[i] = 0;
    for (int k = 0; k < m; k++) {
        const int c = pad(order[m - 1 - k]);
        el[k] = (uint16_t)c;
        out->w[c] = gw::empty((uint32_t)c, (uint32_t)k);
    }
    out->n_el = (int16_t)m;
    out->stones = (int16_t)(p.num_black + p.num_white);
    out->ko_key_col = (int16_t)(p.ko ? p.ko_col : 0);
    out->ko_idx = (int16_t)(p.ko ? pad(p.ko_idx) : 0);
    out->last[0] = (int16_t)(p.last[0] >= 0 ? pad(p.last[0]) : -1);
    out->last[1] = (int16_t)(p.last[1] >= 0 ? pad(p.last[1]) : -1);
    out->ending = p.ending;
    out->pad = 0;
}


// T^steps of TinyMT32 as a nibble table (see JumpTable): column b of T^steps is the state reached
// from the unit vector e_b after `steps` transitions (the transition is linear over GF(2)).
int jump_steps() { return kJumpSteps; }
void build_jump(const RngParams& prm, int steps, JumpTable* out) {
    uint32_t col[128][4];
    for (int b = 0; b < 128; b++) {
        TinyMT t;
        t.s0 = t.s1 = t.s2 = t.s3 = 0;
        (b < 32 ? t.s0 : (b < 64 ? t.s1 : (b < 96 ? t.s2 : t.s3))) = 1u << (b & 31);
        for (int i = 0; i < steps; i++) t.next_state(prm.mat1, prm.mat2);
        col[b][0] = t.s0; col[b][1] = t.s1; col[b][2] = t.s2; col[b][3] = t.s3;
    }
    for (int l = 0; l < 32; l++)
        for (int v = 0; v < 16; v++) {
            uint32_t acc[4] = {0, 0, 0, 0};
            for (int k = 0; k < 4; k++)
                if (v & (1 << k))
                    for (int wd = 0; wd < 4; wd++) acc[wd] ^= col[4 * l + k][wd];
            out->e[l * 16 + v] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
        }
}

// The per-lane matrices of the group kernel (JumpBytes): lane 0 gets T^(lanes*seg), lane j T^(seg*j).
void build_jump_bytes(const RngParams& prm, int lanes, int seg, uint4* out) {
    for (int j = 0; j < lanes; j++) {
        const int steps = j == 0 ? lanes * seg : seg * j;
        uint32_t col[128][4];
        for (int b = 0; b < 128; b++) {
            TinyMT t;
            t.s0 = t.s1 = t.s2 = t.s3 = 0;
            (b < 32 ? t.s0 : (b < 64 ? t.s1 : (b < 96 ? t.s2 : t.s3))) = 1u << (b & 31);
            for (int i = 0; i < steps; i++) t.next_state(prm.mat1, prm.mat2);
            col[b][0] = t.s0; col[b][1] = t.s1; col[b][2] = t.s2; col[b][3] = t.s3;
        }
        for (int by = 0; by < 16; by++)
            for (int v = 0; v < 256; v++) {
                uint32_t acc[4] = {0, 0, 0, 0};
                for (int k = 0; k < 8; k++)
                    if (v & (1 << k))
                        for (int wd = 0; wd < 4; wd++) acc[wd] ^= col[8 * by + k][wd];
                out[((size_t)j * 16 + by) * 256 + v] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
            }
    }
}

}  // namespace agb
