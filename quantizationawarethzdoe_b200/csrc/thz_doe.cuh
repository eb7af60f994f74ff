// Per-pixel bodies of the DOE kernels: stand-alone phase modulation (forward / adjoint) and the
// quantized level selection with its straight-through / surrogate / soft-quantization gradients.
// Host-replayable like everything else (tests/emul).
//
// Reference: Components/QuantizedDOE.py:46-126 (modulate), :1239-1253 + :1379-1388 (STE),
// :1193-1207 (PSQ), :794-860 (score-Gumbel v3), :1022-1031 (naive Gumbel);
// Components/quantization.py:59-122 + utils/Helper_Functions.py:390-398 (nearest-neighbour search
// and its poly / sigmoid surrogate gradients).
#pragma once
#include "thz_asm.cuh"

#define THZ_MAX_LEVELS 64

THZ_HD float thz_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// torch.remainder(a, b) for b > 0 (python-style modulo), then the reference's wrap (a+pi)%(2pi)-pi
THZ_HD float thz_wrap_pi(float a) {
    const float PI_F = 3.14159265358979323846f;
    const float TWO_PI_F = 2.0f * PI_F;
    float r = fmodf(thz_add_rn(a, PI_F), TWO_PI_F);
    if (r < 0.0f) r = thz_add_rn(r, TWO_PI_F);
    return thz_sub_rn(r, PI_F);
}

// ---------------------------------------------------------------- height from weights
// h = hmax * sigmoid(clamp(w, -c, c))   (QuantizedDOE.py:277, :1381 with c = 8; :823 with c = 10)
THZ_HD float thz_height_from_weight(float w, float hmax, float clampv) {
    const float wc = fminf(fmaxf(w, -clampv), clampv);
    return thz_mul_rn(hmax, thz_sigmoid(wc));
}
// d h / d w (torch.clamp passes the gradient on the closed interval)
THZ_HD float thz_height_grad(float w, float hmax, float clampv) {
    if (w < -clampv || w > clampv) return 0.0f;
    const float s = thz_sigmoid(w);
    return hmax * s * (1.0f - s);
}

// ---------------------------------------------------------------- STE: argmin_j |h - lut_j|, first minimum wins
THZ_HD int thz_ste_index(float h, const float* lut, int L) {
    int best = 0;
    float bd = fabsf(thz_sub_rn(h, lut[0]));
    for (int j = 1; j < L; ++j) {
        const float dj = fabsf(thz_sub_rn(h, lut[j]));
        if (dj < bd) {
            bd = dj;
            best = j;
        }
    }
    return best;
}

// ---------------------------------------------------------------- nearest-neighbour search (bucketize right=True, % len(mid))
THZ_HD int thz_nn_index(float x, const float* mid, int nmid) {
    int cnt = 0;
    for (int j = 0; j < nmid; ++j) cnt += (mid[j] <= x) ? 1 : 0;
    return cnt % nmid;
}

// surrogate gradient factor d q / d x of quantization.py:80-96 (kind 1, poly) and :105-122 (kind 2, sigmoid);
// kind 0 is the plain straight-through 1.  `nlut` = len(lut) for python-style negative indexing.
THZ_HD float thz_nn_grad_factor(float x, float q, int idx, const float* lut, int nlut, float s, int kind) {
    if (kind == 0) return 1.0f;
    const float dx = x - q;
    int d = dx > 0.f ? 1 : (dx < 0.f ? -1 : 0);   // reference: NaN.int() -> IndexError when x == q; we take 0
    int oi = idx + d;
    if (oi < 0) oi += nlut;                        // python negative index
    if (oi >= nlut) oi = nlut - 1;                 // reference would raise; clamp instead
    const float other = lut[oi];
    const float mid = (other + q) / 2.0f;
    const float gap = fabsf(other - q) + 1e-20f;
    float z = (x - mid) / gap * 2.0f;
    if (kind == 1) {
        const float base = 1.0f - fabsf(z);
        float v = 0.5f * s * powf(base, s - 1.0f);
        if (v != v) v = 0.0f;                      // nan_to_num
        return v * 2.0f;
    }
    z *= s;
    const float sg = thz_sigmoid(z);
    return sg * (1.0f - sg) * (4.0f * s);
}

// ---------------------------------------------------------------- progressive sigmoid quantisation
// out = delta * sum_{l < L-1} sigmoid(tau (h/delta - 0.5 - l)),  delta = hmax / (L-1)
THZ_HD void thz_psq(float w, float hmax, int L, float tau, float* out, float* dout_dw) {
    const float h = thz_height_from_weight(w, hmax, 8.0f);
    const float delta = hmax / (float)(L - 1);
    const float xn = h / delta - 0.5f;
    float acc = 0.f, dacc = 0.f;
    for (int l = 0; l < L - 1; ++l) {
        const float sg = thz_sigmoid(tau * (xn - (float)l));
        acc += sg;
        dacc += tau * sg * (1.0f - sg);
    }
    *out = 0.0f + delta * acc;
    if (dout_dw) *dout_dw = dacc * thz_height_grad(w, hmax, 8.0f);   // delta * (1/delta) cancels
}

// ---------------------------------------------------------------- Gumbel-softmax (hard=True) over L logits
// y = softmax((score + G) / tau); idx = argmax y (first maximum);
// q = sum_j lut_j * ((onehot_j - y_j) + y_j)   [the forward value of y_hard - y_soft.detach() + y_soft]
// soft_mean = sum_j lut_j y_j                  [what the gradient sees]
THZ_HD int thz_gumbel_hard(const float* logit, const float* lut, int L, float* y, float* q, float* soft_mean) {
    float m = logit[0];
    for (int j = 1; j < L; ++j) m = fmaxf(m, logit[j]);
    float sum = 0.f;
    for (int j = 0; j < L; ++j) {
        y[j] = expf(logit[j] - m);
        sum += y[j];
    }
    int idx = 0;
    float sm = 0.f;
    for (int j = 0; j < L; ++j) {
        y[j] = y[j] / sum;
        if (y[j] > y[idx]) idx = j;
        sm += lut[j] * y[j];
    }
    float qq = 0.f;
    for (int j = 0; j < L; ++j) {
        const float hard = (j == idx) ? 1.0f : 0.0f;
        qq += lut[j] * ((hard - y[j]) + y[j]);
    }
    *q = qq;
    *soft_mean = sm;
    return idx;
}

struct GumbelV3Params {
    float hmax;       // height_constraint_max
    float kfac;       // 2 pi / lambda_min * (sqrt(eps) - 1)   (QuantizedDOE.py:40-41)
    float c_s;
    float tau;
    float tau_max;
    float s;          // float32(tau_max / tau)
    float beta;       // blend weight: h_out = (1-beta) h + beta q;  beta >= 1 -> pure q (iter_frac > 0.8)
    float omb;        // float32(1 - beta) rounded from the python double, as the reference multiplies it
    int L;
    int phase_input;  // 1: `w` already is the phase (SoftGumbelQuantizedDOELayer v1, QuantizedDOE.py:436-446)
};

// SoftGumbelQuantizedDOELayerv3.preprocessed_height_map for one pixel (iter_frac > 0.3 branch).
//   noise: G[j] for this pixel at stride `nstride`.  Returns idx; *h_out, and *dh_dw (if non-null).
THZ_HD int thz_gumbel_v3_pixel(float w, const float* lut, const float* noise, size_t nstride, const GumbelV3Params& P,
                               float* h_out, float* dh_dw) {
    const float PI_F = 3.14159265358979323846f;
    const float h = P.phase_input ? 0.0f : thz_height_from_weight(w, P.hmax, 10.0f);
    const float phase = P.phase_input ? w : thz_mul_rn(P.kfac, h);
    const float dphase_dh = P.phase_input ? 1.0f : P.kfac;
    const float wp = thz_wrap_pi(phase);
    const float s = P.s;
    float logit[THZ_MAX_LEVELS], y[THZ_MAX_LEVELS], dsc[THZ_MAX_LEVELS];
    for (int j = 0; j < P.L; ++j) {
        const float pl = thz_wrap_pi(thz_mul_rn(P.kfac, lut[j]));
        float diff = thz_wrap_pi(thz_sub_rn(wp, pl));
        diff = diff / PI_F;
        const float z = s * diff;
        const float sg = thz_sigmoid(z);
        const float score = sg * (1.0f - sg) * 4.0f * P.c_s * s;
        logit[j] = (score + noise[(size_t)j * nstride]) / P.tau;
        // d score / d h = 4 c_s s * sg(1-sg)(1-2sg) * s * kfac / pi
        dsc[j] = 4.0f * P.c_s * s * sg * (1.0f - sg) * (1.0f - 2.0f * sg) * s * dphase_dh / PI_F;
    }
    float q, sm;
    const int idx = thz_gumbel_hard(logit, lut, P.L, y, &q, &sm);
    *h_out = (P.beta >= 1.0f || P.phase_input) ? q : thz_add_rn(thz_mul_rn(P.omb, h), thz_mul_rn(P.beta, q));
    if (dh_dw) {
        float dq_dh = 0.f;
        for (int j = 0; j < P.L; ++j) dq_dh += y[j] * (lut[j] - sm) * dsc[j] / P.tau;
        if (P.phase_input) {
            *dh_dw = dq_dh;
        } else {
            const float dout_dh = (P.beta >= 1.0f) ? dq_dh : P.omb + P.beta * dq_dh;
            *dh_dw = dout_dh * thz_height_grad(w, P.hmax, 10.0f);
        }
    }
    return idx;
}

// ---------------------------------------------------------------- thickness-space softmax quantization
// SoftmaxBasedQuantization.forward + score_thickness('sigmoid') (Components/quantization.py:36-46, 128-161) for one pixel:
//   nd_j = (t - lut_j) / m,  m = max over the WHOLE map and all levels of |t - lut_j|   (score_thickness :40-41)
//   score_j = sigmoid(s nd_j) (1 - sigmoid(s nd_j)) 4 * c * s,   s = tau_max / tau
//   gumbel: y = softmax((score + G) / tau), hard -> (onehot - y) + y      [F.gumbel_softmax]
//   plain:  y = softmax(score / tau),       hard -> (onehot + y) - y      [:146-152]
//   q = sum_j onehot_j lut_j
// For the backward (autograd through the out-of-place form diff / max|diff|; the reference's in-place `diff /= max` makes
// its own backward raise) the pixel also returns
//   A = d q / d t at fixed m,   Bm = d q / d m,   E = sum over the levels of this pixel that attain the maximum of sign(diff)
// so that  gt = g A + (sum_pixels g Bm) / ties * E  -- torch.max() splits its gradient evenly over ties.
struct SoftmaxQParams {
    float m;          // global max |t - lut_j|
    float s, c, tau;
    int L, hard, gumbel;
};
THZ_HD int thz_softmaxq_pixel(float t, const float* lut, const float* noise, size_t nstride, const SoftmaxQParams& P, float* q,
                              float* A, float* Bm, float* E, int* ties) {
    float logit[THZ_MAX_LEVELS], y[THZ_MAX_LEVELS], dsc[THZ_MAX_LEVELS], diff[THZ_MAX_LEVELS];
    float e = 0.f;
    int nt = 0;
    for (int j = 0; j < P.L; ++j) {
        diff[j] = thz_sub_rn(t, lut[j]);
        if (fabsf(diff[j]) == P.m) {
            e += diff[j] > 0.f ? 1.f : (diff[j] < 0.f ? -1.f : 0.f);
            ++nt;
        }
        const float nd = diff[j] / P.m;
        const float z = thz_mul_rn(P.s, nd);
        const float sg = thz_sigmoid(z);
        const float score = thz_mul_rn(thz_mul_rn(thz_mul_rn(thz_mul_rn(sg, thz_sub_rn(1.0f, sg)), 4.0f), P.c), P.s);
        logit[j] = P.gumbel ? thz_add_rn(score, noise[(size_t)j * nstride]) / P.tau : score / P.tau;
        dsc[j] = 4.0f * P.c * P.s * sg * (1.0f - sg) * (1.0f - 2.0f * sg) * P.s;        // d score / d nd
    }
    float mx = logit[0];
    for (int j = 1; j < P.L; ++j) mx = fmaxf(mx, logit[j]);
    float sum = 0.f;
    for (int j = 0; j < P.L; ++j) {
        y[j] = expf(logit[j] - mx);
        sum += y[j];
    }
    int idx = 0;
    float sm = 0.f;
    for (int j = 0; j < P.L; ++j) {
        y[j] = y[j] / sum;
        if (y[j] > y[idx]) idx = j;
        sm += lut[j] * y[j];
    }
    float qq = 0.f;
    for (int j = 0; j < P.L; ++j) {
        float oh = y[j];
        if (P.hard) {
            const float hd = (j == idx) ? 1.0f : 0.0f;
            oh = P.gumbel ? thz_add_rn(thz_sub_rn(hd, y[j]), y[j]) : thz_sub_rn(thz_add_rn(hd, y[j]), y[j]);
        }
        qq = thz_add_rn(qq, thz_mul_rn(oh, lut[j]));
    }
    *q = qq;
    if (A) {
        float a = 0.f, b = 0.f;
        for (int j = 0; j < P.L; ++j) {
            const float dq_dnd = y[j] * (lut[j] - sm) * dsc[j] / P.tau;
            a += dq_dnd;
            b -= dq_dnd * diff[j];
        }
        *A = a / P.m;
        *Bm = b / (P.m * P.m);
        *E = e;
    }
    *ties = nt;
    return idx;
}

// score_thickness for every scoring function the reference offers (quantization.py:36-55): 0 sigmoid, 1 log, 2 poly, 3 sine,
// 4 chirp; nd = diff / max|diff| already formed.
THZ_HD float thz_score_value(float nd, float s, int func) {
    const float PI_F = 3.14159265358979323846f;
    switch (func) {
    case 0: {
        const float sg = thz_sigmoid(thz_mul_rn(s, nd));
        return thz_mul_rn(thz_mul_rn(sg, thz_sub_rn(1.0f, sg)), 4.0f);
    }
    case 1: return thz_mul_rn(-logf(thz_add_rn(fabsf(nd), 1e-20f)), s);
    case 2: return thz_sub_rn(1.0f, powf(fabsf(nd), s));
    case 3: return cosf(thz_mul_rn(PI_F, fminf(fmaxf(thz_mul_rn(s, nd), -1.0f), 1.0f)));
    default: return thz_sub_rn(1.0f, cosf(thz_mul_rn(PI_F, powf(thz_sub_rn(1.0f, fabsf(nd)), s))));
    }
}
