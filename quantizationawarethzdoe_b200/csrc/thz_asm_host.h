// Host-side launch planning for the fused ASM pipeline, shared by the CUDA launcher (thz_asm.cu)
// and the CPU replay harness (tests/emul).  No CUDA runtime calls in here.
#pragma once
#include <stdlib.h>
#include "thz_asm_p2.cuh"

#define THZ_SMEM_BUDGET (200 * 1024)   // per-CTA dynamic shared memory we allow ourselves (HW max 227 KB)

struct AsmLaunch {
    RowFwdArgs k1;
    ColArgs k2;
    RowInvArgs k3;
    int k1_grid, k1_threads;
    size_t k1_smem;
    int k2_gridx, k2_threads;
    size_t k2_smem;
    int k3_gridx, k3_gridy, k3_threads;
    size_t k3_smem;
    int mixed_w, mixed_h;
    int p2_w, p2_h;           // 1: the compile-time specialised power-of-two kernels serve this axis
    int gh_needs_zero;        // 1: gh is accumulated with atomics and must be zeroed first
};

static inline int thz_imax(int a, int b) { return a > b ? a : b; }
static inline int thz_imin(int a, int b) { return a < b ? a : b; }

static inline int thz_asm_validate(const thz_asm_desc* d) {
    if (!d) return THZ_E_NULL;
    const int st = d->stages ? d->stages : 7;
    if (st < 0 || st > 7) return THZ_E_SHAPE;
    if (st != 7 && d->bc_chunk > 0 && d->bc_chunk < d->B * d->C) return THZ_E_SHAPE;   // staged runs keep all fields in ws
    if (((st & 1) && !d->x) || !d->tw_h || !d->tw_w || (!d->ws && (d->slab_parts <= 1 || st == 2))) return THZ_E_NULL;
    if (d->B < 1 || d->C < 1 || d->inH < 1 || d->inW < 1 || d->outH < 1 || d->outW < 1) return THZ_E_SHAPE;
    if (d->in_r0 < 0 || d->in_c0 < 0 || d->out_r0 < 0 || d->out_c0 < 0) return THZ_E_SHAPE;
    if (d->in_r0 + d->inH > d->Hp || d->in_c0 + d->inW > d->Wp) return THZ_E_SHAPE;
    if (d->out_r0 + d->outH > d->Hp || d->out_c0 + d->outW > d->Wp) return THZ_E_SHAPE;
    if ((st & 2) && d->tf_mode == 0 && (!d->tf_rowvec || !d->tf_colvec || !d->tf_scal)) return THZ_E_NULL;
    if ((st & 2) && d->tf_mode == 1 && !d->tf_table) return THZ_E_NULL;
    if (d->tf_mode < 0 || d->tf_mode > 2) return THZ_E_SHAPE;
    if (d->doe_mode < 0 || d->doe_mode > 2) return THZ_E_SHAPE;
    if (d->doe_mode != 0 && (!d->doe_hmap || !d->doe_coef)) return THZ_E_NULL;
    if ((st & 4) && d->doe_mode == 2 && (!d->doe_xsaved || !d->doe_gh)) return THZ_E_NULL;
    if (d->elem_mode < 0 || d->elem_mode > 2 || (d->elem_mode != 0 && !d->elem_mask && !d->elem_mul)) return THZ_E_SHAPE;
    if (d->elem_mode != 0 && d->slab_parts > 1) return THZ_E_UNSUPPORTED;      // the slab pipeline fuses the DOE only
    if (d->doe_gh_mode < 0 || d->doe_gh_mode > 1) return THZ_E_SHAPE;
    if (d->doe_levels < 0 || (d->doe_levels > 0 && (!d->doe_level_idx) != (!d->doe_level_phase))) return THZ_E_SHAPE;
    if (d->doe_hmap_bstride < 0 || (d->doe_hmap_bstride != 0 && d->doe_mode != 1)) return THZ_E_SHAPE;   // per-entry maps: forward only
    if ((st & 4) && d->doe_mode != 2 && !d->y) return THZ_E_NULL;
    if (d->slab_parts > 1) {
        if (d->slab_parts > 8 || (st != 1 && st != 2 && st != 4) || d->slab_rows < 1 || d->slab_row0 < 0) return THZ_E_SHAPE;
        if (st == 2) {          // local column pass between the two slabs: ws = S1 (blocked), slab_ptrs[0] = S2 (row-major)
            if ((d->slab_blocked && d->Wp % 4) || d->slab_rows < thz_imax(d->inH, d->outH)) return THZ_E_SHAPE;
            if (!d->ws || !d->slab_ptrs[0]) return THZ_E_NULL;
        } else {
            if (d->Wp % d->slab_parts || (d->slab_blocked && (d->Wp / d->slab_parts) % 4)) return THZ_E_SHAPE;
            if (d->slab_row0 + (st == 1 ? d->inH : d->outH) > d->slab_rows) return THZ_E_SHAPE;
            for (int i = 0; i < d->slab_parts; ++i)
                if (!d->slab_ptrs[i]) return THZ_E_NULL;
        }
    }
    return THZ_OK;
}

static inline uint64_t thz_asm_chunk_fields(const thz_asm_desc* d) {
    uint64_t nbc = (uint64_t)d->B * d->C;
    if (d->bc_chunk > 0 && (uint64_t)d->bc_chunk < nbc) return (uint64_t)d->bc_chunk;
    return nbc;
}

static inline uint64_t rowsT_ws(const thz_asm_desc* d) { return (uint64_t)thz_imax(d->inH, d->outH); }
// Lengths the static kernels serve and THZ_NO_P2 / THZ_NO_TILED do not veto: such whole-pipeline runs keep TWO
// intermediates (blocked row spectra, row-major column-pass output), everything else one.
static inline bool thz_asm_two_buffers(const thz_asm_desc* d);
static inline uint64_t thz_asm_ws_bytes(const thz_asm_desc* d) {
    const uint64_t one = thz_asm_chunk_fields(d) * rowsT_ws(d) * (uint64_t)d->Wp * sizeof(cpx);
    return thz_asm_two_buffers(d) ? 2 * one : one;
}


static inline bool thz_is_p2_size(int n) { return thz_sp_instantiated(n); }
static inline bool thz_env_is_1(const char* name) {
    const char* v = getenv(name);
    return v && v[0] == '1';
}
static inline bool thz_asm_two_buffers(const thz_asm_desc* d) {
    const int st = d->stages ? d->stages : 7;
    return st == 7 && d->slab_parts <= 1 && d->Wp % 4 == 0 && thz_is_p2_size(d->Wp) && thz_is_p2_size(d->Hp) &&
           !thz_env_is_1("THZ_NO_P2") && !thz_env_is_1("THZ_NO_TILED");
}

static inline bool p2_k2_fast_ok_rt(int n) {
    if (!thz_sp_instantiated(n) || n < 256) return false;
#define THZ_SP_FASTCMP(NN) if (n == NN) return sp_k2_fast_ok_or_false<NN>();
    THZ_SP_SIZES(THZ_SP_FASTCMP)
#undef THZ_SP_FASTCMP
    return false;
}
static inline int thz_p2_min_blocks_rt(int n) {
#define THZ_SP_MB(NN) if (n == NN) return p2_min_blocks(NN);
    THZ_SP_SIZES(THZ_SP_MB)
#undef THZ_SP_MB
    return 3;
}
static inline int thz_p2_row_lines_rt(int n) { return p2_row_lines(n); }
static inline int thz_p2_tw_count_rt(int n) { return p2_tw_count(n); }
static inline int thz_p2_row_threads_rt(int n) { return p2_row_threads(n); }
static inline int thz_p2_col_cols_rt(int n) { return p2_col_cols(n); }
static inline int thz_p2_col_threads_rt(int n) { return p2_col_threads(n); }

// Switch the launch geometry of the axes that the power-of-two fast path serves (THZ_NO_P2=1 disables it).
static inline void thz_asm_apply_p2(const thz_asm_desc* d, int nbc, int sm_count, AsmLaunch* L) {
    const char* off = getenv("THZ_NO_P2");
    const bool enabled = !(off && off[0] == '1');
    L->p2_w = enabled && thz_is_p2_size(d->Wp);
    L->p2_h = enabled && thz_is_p2_size(d->Hp);
    if (L->p2_w) {
        const int lines = thz_p2_row_lines_rt(d->Wp), threads = thz_p2_row_threads_rt(d->Wp);
        const size_t smem = (size_t)lines * (d->Wp + (d->Wp >> 4)) * sizeof(cpx);
        L->k1.lines = lines;
        L->k1_threads = threads;
        L->k1_grid = (nbc * d->inH + lines - 1) / lines;
        // line buffer + staged raw rows (+ staged height-map rows), see thz_p2_k1
        const size_t tw_bytes = (size_t)thz_p2_tw_count_rt(d->Wp) * sizeof(cpx);   // shared-memory twiddle copy
        L->k1_smem = smem + tw_bytes;
        if (p2_row_pipelined(d->Wp)) L->k1_smem += (size_t)lines * d->inW * sizeof(cpx) + (size_t)lines * d->inW * sizeof(float);
        L->k3.lines = lines;
        L->k3_threads = threads;
        L->k3_gridx = (d->outH + lines - 1) / lines;
        // Split of the field axis over gridDim.y.  Each CTA walks bc_per_cta fields (register accumulators for grad_height, the
        // next field's rows prefetched), so the cost of a launch is  waves x bc_per_cta  with waves counted against the CTAs
        // that are resident at once (3 per SM, 1 for the 8192+ lines): pick the split that minimises it.  The old rule
        // (double until the grid reaches 4 x SMs) gave 640 CTAs on 444 slots for the 400-point lines of the DONN config --
        // 1.44 waves, the second one less than half full (profiles/r02_donn_ncu_summary.txt).
        const long resident = (long)sm_count * thz_p2_min_blocks_rt(d->Wp);
        int gy = 1;
        if (L->k3_gridx < resident && nbc > 1) {
            long best = -1;
            const int gy_max = thz_imin(nbc, (int)(4 * resident / L->k3_gridx) + 1);
            for (int cand = 1; cand <= gy_max; ++cand) {
                const int per = (nbc + cand - 1) / cand, eff = (nbc + per - 1) / per;
                const long waves = ((long)L->k3_gridx * eff + resident - 1) / resident;
                const long cost = waves * per;
                if (best < 0 || cost < best) {
                    best = cost;
                    gy = eff;
                }
            }
        }
        L->k3.bc_per_cta = (nbc + gy - 1) / gy;
        L->k3_gridy = (nbc + L->k3.bc_per_cta - 1) / L->k3.bc_per_cta;
        L->k3_smem = (p2_row_pipelined(d->Wp) ? 2 * smem : smem) + tw_bytes;   // double-buffered line (thz_p2_k3 prefetches the next field)
    }
    if (L->p2_h) {
        const int cols = thz_p2_col_cols_rt(d->Hp);
        L->k2.cols = cols;
        L->k2_gridx = (d->Wp + cols - 1) / cols;
        L->k2_threads = thz_p2_col_threads_rt(d->Hp);
        L->k2_smem = (size_t)cols * (d->Hp + (d->Hp >> 4)) * sizeof(cpx) + (size_t)thz_p2_tw_count_rt(d->Hp) * sizeof(cpx);
    }
}

// Build the three kernels' arguments for the chunk of fields [f0, f0+nbc).
static inline int thz_asm_plan_chunk(const thz_asm_desc* d, int f0, int nbc, int sm_count, AsmLaunch* L) {
    FftPlan pw, ph;
    if (thz_make_plan(d->Wp, &pw) != 0) return THZ_E_UNSUPPORTED;
    if (thz_make_plan(d->Hp, &ph) != 0) return THZ_E_UNSUPPORTED;
    L->mixed_w = pw.mixed;
    L->mixed_h = ph.mixed;
    const int rowsT = thz_imax(d->inH, d->outH);
    const size_t line_bytes_w = (size_t)thz_padded_len(d->Wp) * sizeof(cpx);
    const size_t line_bytes_h = (size_t)thz_padded_len(d->Hp) * sizeof(cpx);
    if (line_bytes_w > THZ_SMEM_BUDGET || line_bytes_h > THZ_SMEM_BUDGET) return THZ_E_SMEM;

    // ---- K1
    RowFwdArgs& a1 = L->k1;
    a1.x = (const cpx*)d->x + (size_t)f0 * d->inH * d->inW;
    a1.T = (cpx*)d->ws;
    a1.nbc = nbc;
    a1.rowsT = rowsT;
    a1.c0 = f0 % d->C;
    a1.C = d->C;
    a1.inH = d->inH;
    a1.inW = d->inW;
    a1.Wp = d->Wp;
    a1.in_c0 = d->in_c0;
    a1.plan = pw;
    a1.tw = (const cpx*)d->tw_w;
    a1.doe.hmap = d->doe_mode == 1 ? (const float*)d->doe_hmap : nullptr;
    a1.doe.coef = (const float4*)d->doe_coef;
    a1.doe.base = d->doe_base;
    a1.doe.b0 = f0 / d->C;
    a1.doe.hstride = d->doe_mode == 1 ? d->doe_hmap_bstride : 0;
    a1.doe.lphase = nullptr;
    a1.doe.nlev = 0;
    a1.elem.mask = d->elem_mode == 1 ? (const float*)d->elem_mask : nullptr;
    a1.elem.mul = d->elem_mode == 1 ? (const cpx*)d->elem_mul : nullptr;
    a1.conj_in = 0;
    memset(&a1.slab, 0, sizeof(a1.slab));
    memset(&L->k3.slab, 0, sizeof(L->k3.slab));
    if (d->slab_parts > 1) {
        SlabArgs sl;
        sl.parts = d->slab_parts;
        sl.row0 = d->slab_row0;
        sl.rows = d->slab_rows;
        sl.Wc = d->Wp / d->slab_parts;
        sl.blocked = 0;
        for (int i = 0; i < 8; ++i) sl.ptr[i] = i < d->slab_parts ? (cpx*)d->slab_ptrs[i] : nullptr;
        const int st_ = d->stages ? d->stages : 7;
        if (st_ != 2) {
            L->k3.slab = sl;          // stage 4 gathers from the row-major slabs S2
            sl.blocked = d->slab_blocked ? 1 : 0;   // stage 1 scatters into S1 (optionally in 4-column blocks)
            a1.slab = sl;
        }
    }
    {
        // enough lines per CTA to give 256 threads at least one radix-16 butterfly each
        int lines = thz_imax(1, 4096 / d->Wp);
        lines = thz_imin(lines, 16);
        while (lines > 1 && lines * line_bytes_w > 64 * 1024) --lines;
        if (d->tune_lines > 0 && d->tune_lines * line_bytes_w <= THZ_SMEM_BUDGET) lines = d->tune_lines;
        a1.lines = lines;
        L->k1_threads = 256;
        L->k1_grid = (nbc * d->inH + lines - 1) / lines;
        L->k1_smem = lines * line_bytes_w;
    }

    // ---- K2
    ColArgs& a2 = L->k2;
    a2.T = (cpx*)d->ws;
    a2.nbc = nbc;
    a2.c0 = f0 % d->C;
    a2.C = d->C;
    a2.rowsT = rowsT;
    a2.inH = d->inH;
    a2.outH = d->outH;
    a2.Hp = d->Hp;
    a2.Wp = d->Wp;
    a2.in_r0 = d->in_r0;
    a2.out_r0 = d->out_r0;
    a2.plan = ph;
    a2.tw = (const cpx*)d->tw_h;
    a2.tf.mode = d->tf_mode;
    a2.tf.conj = d->tf_conj;
    a2.tf.rowvec = (const float2*)d->tf_rowvec;
    a2.tf.colvec = (const float*)d->tf_colvec;
    a2.tf.scal = (const float2*)d->tf_scal;
    a2.tf.table = (const cpx*)d->tf_table;
    a2.tf.row_chunked = d->tf_row_chunked;
    {
        int cols = 16;
        // shrink the tile until it fits and until there are enough tiles to fill the GPU twice
        while (cols > 1 && cols * line_bytes_h > 72 * 1024) cols >>= 1;
        while (cols > 4 && (long)((d->Wp + cols - 1) / cols) * nbc < 2L * sm_count) cols >>= 1;
        if (d->tune_k2_cols > 0) cols = d->tune_k2_cols;
        if ((size_t)cols * line_bytes_h > THZ_SMEM_BUDGET) return THZ_E_SMEM;
        a2.cols = cols;
        L->k2_gridx = (d->Wp + cols - 1) / cols;
        L->k2_smem = cols * line_bytes_h;
        const int work = (d->Hp / 16 + 1) * cols;
        L->k2_threads = work >= 1024 ? 512 : (work >= 384 ? 256 : 128);
    }

    // ---- K3
    RowInvArgs& a3 = L->k3;
    a3.T = (const cpx*)d->ws;
    a3.y = d->y ? (cpx*)d->y + (size_t)f0 * d->outH * d->outW : nullptr;
    a3.nbc = nbc;
    a3.c0 = f0 % d->C;
    a3.C = d->C;
    a3.rowsT = rowsT;
    a3.outH = d->outH;
    a3.outW = d->outW;
    a3.Wp = d->Wp;
    a3.out_c0 = d->out_c0;
    a3.scale = 1.0f / ((float)d->Hp * (float)d->Wp);
    a3.plan = pw;
    a3.tw = (const cpx*)d->tw_w;
    a3.doe.hmap = d->doe_mode == 2 ? (const float*)d->doe_hmap : nullptr;
    a3.doe.coef = (const float4*)d->doe_coef;
    a3.doe.base = d->doe_base;
    a3.doe.b0 = 0;
    a3.doe.hstride = 0;
    a3.doe.lphase = nullptr;
    a3.doe.nlev = 0;
    a3.elem.mask = d->elem_mode == 2 ? (const float*)d->elem_mask : nullptr;
    a3.elem.mul = d->elem_mode == 2 ? (const cpx*)d->elem_mul : nullptr;
    a3.xsaved = d->doe_mode == 2 ? (const cpx*)d->doe_xsaved + (size_t)f0 * d->outH * d->outW : nullptr;
    a3.gh = (float*)d->doe_gh;
    {
        int lines = thz_imax(1, 4096 / d->Wp);
        lines = thz_imin(lines, 16);
        while (lines > 1 && lines * line_bytes_w > 64 * 1024) --lines;
        int threads = 256;
        if (d->tune_lines > 0 && d->tune_lines * line_bytes_w <= THZ_SMEM_BUDGET) lines = d->tune_lines;
        if (d->doe_mode == 2) {
            // register accumulators: lines*outW <= THZ_K3_OWN * threads
            while (lines > 1 && (long)lines * d->outW > (long)THZ_K3_OWN * threads) --lines;
            while (threads < 512 && (long)lines * d->outW > (long)THZ_K3_OWN * threads) threads <<= 1;
            if ((long)lines * d->outW > (long)THZ_K3_OWN * threads) return THZ_E_UNSUPPORTED;
        }
        a3.lines = lines;
        L->k3_threads = threads;
        L->k3_gridx = (d->outH + lines - 1) / lines;
        // split the field axis over gridDim.y until the grid fills the GPU ~4x
        int gy = 1;
        while (gy < nbc && (long)L->k3_gridx * gy < 4L * sm_count) gy <<= 1;
        gy = thz_imin(gy, nbc);
        a3.bc_per_cta = (nbc + gy - 1) / gy;
        L->k3_gridy = (nbc + a3.bc_per_cta - 1) / a3.bc_per_cta;
        L->k3_smem = lines * line_bytes_w;
    }
    thz_asm_apply_p2(d, nbc, sm_count, L);
    // quantised DOE through the static row kernels: the level-index map takes the place of the height map and the transmission
    // of every level comes from a [C][levels] table (thz_asm_desc.doe_level_*; THZ_NO_DOE_LUT=1: evaluate per pixel as before)
    if (L->p2_w && d->doe_mode != 0 && d->doe_level_idx && d->doe_level_phase && d->doe_levels > 0 && d->elem_mode == 0 &&
        d->doe_hmap_bstride == 0 && d->slab_parts <= 1 && !thz_env_is_1("THZ_NO_DOE_LUT")) {
        DoeArgs& dd = d->doe_mode == 1 ? L->k1.doe : L->k3.doe;
        dd.hmap = (const float*)d->doe_level_idx;
        dd.lphase = (const cpx*)d->doe_level_phase;
        dd.nlev = d->doe_levels;
    }
    {
        // blocked intermediate between the static kernels of a whole-pipeline run (THZ_NO_TILED=1 keeps row-major T)
        int tiled = thz_asm_two_buffers(d) ? 2 : 0, tiled2 = 0;
        if (tiled) {      // experiment knobs: log2 of the block widths of the two intermediates
            const char* e1 = getenv("THZ_T1_LOG2");
            const char* e2 = getenv("THZ_T2_LOG2");
            if (e1 && e1[0] >= '1' && e1[0] <= '4' && d->Wp % (1 << (e1[0] - '0')) == 0) tiled = e1[0] - '0';
            if (e2 && e2[0] >= '1' && e2[0] <= '4' && d->Wp % (1 << (e2[0] - '0')) == 0) tiled2 = e2[0] - '0';
        }
        L->k1.t_tiled = tiled;
        L->k2.t_tiled = tiled;
        L->k2.tout_tiled = tiled2;
        L->k3.t_tiled = tiled2;
        L->k2.Tout = nullptr;
        if ((d->stages ? d->stages : 7) == 2 && d->slab_parts > 1) {   // column pass of a peer-memory slab FFT: S1 (ws) -> S2
            L->k2.t_tiled = d->slab_blocked ? 2 : 0;
            L->k2.tout_tiled = 0;
            L->k2.Tout = (cpx*)d->slab_ptrs[0];
            if (!L->p2_h) return THZ_E_UNSUPPORTED;
        }
        if (tiled) {     // second half of the workspace: the column kernel's row-major output, read by the row-iFFT kernel
            L->k2.Tout = (cpx*)d->ws + (size_t)thz_asm_chunk_fields(d) * rowsT_ws(d) * d->Wp;
            L->k3.T = L->k2.Tout;
        }
    }
    {   // centred 2x padding / crop: the static kernels prune their first / last radix-16 stage (THZ_NO_PRUNE=1: off)
        const bool on = !thz_env_is_1("THZ_NO_PRUNE");
        L->k1.half_in = (on && 4 * d->in_c0 == d->Wp && 2 * d->inW == d->Wp) ? 1 : 0;
        L->k2.half_in = (on && 4 * d->in_r0 == d->Hp && 2 * d->inH == d->Hp) ? 1 : 0;
        L->k2.half_out = (on && 4 * d->out_r0 == d->Hp && 2 * d->outH == d->Hp) ? 1 : 0;
        L->k3.half_out = (on && 4 * d->out_c0 == d->Wp && 2 * d->outW == d->Wp) ? 1 : 0;
    }
    // the specialised column kernel (thz_p2_k2f): centred 2x padding, blocked input, row-major output, whole tiles, static
    // length with a radix-16 first stage, transfer function from chunked vectors / a table / none (THZ_NO_K2FAST=1: off)
    // look-ahead of the column kernels' L2 prefetch: the tile of the CTA that takes this one's place = one residency further on
    // in launch order (THZ_K2_PF: multiple of that distance in percent, 0 = off)
    L->k2.pf_blocks = 0;
    if (L->p2_h && L->k2.t_tiled == 2 && d->Wp % 4 == 0 && d->Wp % L->k2.cols == 0 && d->slab_parts <= 1) {
        const char* e = getenv("THZ_K2_PF");
        const long pct = e ? atol(e) : 100;
        const long resident = (long)sm_count * thz_p2_min_blocks_rt(d->Hp);
        const long blocks = (resident * L->k2.cols * pct) / 400;
        L->k2.pf_blocks = (int)(pct <= 0 ? 0 : (blocks < 1 ? 1 : blocks));
    }
    L->k3.pf_groups = 0;
    L->k2.t2_perm = 0;        // thz_asm_propagate turns the permuted intermediate on (device library only)
    L->k3.t2_perm = 0;
    if (L->p2_w && L->k3.t_tiled == 0 && d->slab_parts <= 1) {
        const char* e = getenv("THZ_K3_PF");
        const long pct = e ? atol(e) : 100;
        const long resident = (long)sm_count * thz_p2_min_blocks_rt(d->Wp);
        L->k3.pf_groups = (int)(pct <= 0 ? 0 : (resident * pct / 100 < 1 ? 1 : resident * pct / 100));
    }
    L->k2.fast = (L->p2_h && L->k2.half_in && L->k2.half_out && L->k2.t_tiled == 2 && L->k2.tout_tiled == 0 && L->k2.Tout &&
                  p2_k2_fast_ok_rt(d->Hp) && d->Wp % L->k2.cols == 0 && d->Wp % 4 == 0 &&
                  (d->tf_mode != 0 || d->tf_row_chunked) && L->k2_threads == thz_p2_col_threads_rt(d->Hp) &&
                  !thz_env_is_1("THZ_NO_K2FAST")) ? 1 : 0;
    if (d->slab_parts > 1 && (d->stages & 5) && !L->p2_w) return THZ_E_UNSUPPORTED;   // only the static row kernels scatter / gather slabs
    if (d->tf_row_chunked && d->tf_mode == 0 && !L->p2_h) return THZ_E_UNSUPPORTED;   // chunked row vectors: static column kernels only
    return THZ_OK;
}

// Radix R > 0 if this chunk can run with the column-permuted K2 -> K3 intermediate and the TMA-staged row-iFFT kernel
// (ColArgs.t2_perm = RowInvArgs.t2_perm = R): whole pipeline on the static kernels, fast column kernel with 2-column tiles,
// row-major second buffer, a line length thz_p2_k3t serves.  Applied by thz_asm_propagate (off: THZ_NO_K3TMA=1) and, for the
// CPU replay of the same index arithmetic, by the test harness (on: THZ_EMUL_T2_PERM=1).
static inline int thz_asm_t2_perm_radix(const thz_asm_desc* d, const AsmLaunch* L, int stages) {
    if (stages != 7 || d->slab_parts > 1 || !L->p2_w || !L->p2_h || !L->k2.fast || L->k2.cols != 2 || L->k2.tout_tiled != 0 ||
        L->k3.t_tiled != 0 || L->k3.T != L->k2.Tout)
        return 0;
    return p2_k3_tma_ok(d->Wp) ? p2_radix(d->Wp, p2_stages(d->Wp) - 1) : 0;
}
