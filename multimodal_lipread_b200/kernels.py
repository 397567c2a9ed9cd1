"""Thin tensor-level wrappers over the C ABI (include/lipread_b200.h).  Every function takes CUDA
tensors (torch is used for device memory and the current stream only), passes raw pointers to
liblipread_b200.so and returns nothing: outputs are caller-allocated, as the ABI requires."""
import torch

from ._lib import lib, check, ACT_NONE, ACT_RELU, ACT_HSWISH, ACT_HSIGMOID, ACT_RELU6  # noqa: F401


def _p(t):
    return None if t is None else t.data_ptr()


def _s():
    return torch.cuda.current_stream().cuda_stream


def gemm(A, lda, a_trans, B, ldb, b_trans, C, ldc, M, N, K, bias=None, act=ACT_NONE, R=None, ldr=0, stats=None,
         ksplit=1):
    check(lib.lr_gemm(_p(A), lda, a_trans, _p(B), ldb, b_trans, _p(C), ldc, M, N, K, _p(bias), act, _p(R), ldr,
                      _p(stats), ksplit, _s()))


def auto_ksplit(M, N, K, sms=148, min_k=256):
    """Split the reduction so that a GEMM with few output tiles still fills the GPU."""
    tiles = ((M + 63) // 64) * ((N + 63) // 64)
    want = max(1, (2 * sms) // tiles)
    return max(1, min(want, K // min_k if K >= min_k else 1))


def linear_fwd(x, w, out, bias=None, act=ACT_NONE, stats=None, M=None, lda=None, ldc=None, ksplit=1):
    """out[M,N] = act(x[M,K] @ w[N,K]^T + bias)."""
    N, K = w.shape[0], w[0].numel()
    M = x.shape[0] if M is None else M
    gemm(x, K if lda is None else lda, 0, w, K, 0, out, N if ldc is None else ldc, M, N, K, bias=bias, act=act,
         stats=stats, ksplit=ksplit)


def linear_dgrad(dy, w, dx, M=None, ldy=None, ldx=None, accumulate=False):
    """dx[M,K] (+)= dy[M,N] @ w[N,K]."""
    N, K = w.shape[0], w[0].numel()
    M = dy.shape[0] if M is None else M
    ldx = K if ldx is None else ldx
    gemm(dy, N if ldy is None else ldy, 0, w, K, 1, dx, ldx, M, K, N, R=dx if accumulate else None, ldr=ldx)


def linear_wgrad(dy, x, dw, M=None, ldy=None, ldx=None, sms=148):
    """dw[N,K] += dy[M,N]^T @ x[M,K]   (dw holds the running gradient; split-K atomics add onto it)."""
    N, K = dw.shape[0], dw[0].numel()
    M = dy.shape[0] if M is None else M
    ks = auto_ksplit(N, K, M, sms)
    if ks > 1:
        gemm(dy, N if ldy is None else ldy, 1, x, K if ldx is None else ldx, 1, dw, K, N, K, M, ksplit=ks)
    else:
        gemm(dy, N if ldy is None else ldy, 1, x, K if ldx is None else ldx, 1, dw, K, N, K, M, R=dw, ldr=K)


def stem_conv_fwd(x, layout, w, y, stats, scale):
    """layout = (is_u8, B, T, H, W, sb, st, sc, sh, sw)."""
    check(lib.lr_stem_conv_fwd(_p(x), *layout, scale, _p(w), _p(y), _p(stats), _s()))


def stem_conv_wgrad(x, layout, dy, dw, scale):
    check(lib.lr_stem_conv_wgrad(_p(x), *layout, scale, _p(dy), _p(dw), _s()))


def dwconv_fwd(x, w, y, stats, F, H, W, C, k, stride):
    check(lib.lr_dwconv_fwd(_p(x), _p(w), _p(y), _p(stats), F, H, W, C, k, stride, _s()))


def dwconv_dgrad(dy, w, dx, F, H, W, C, k, stride):
    check(lib.lr_dwconv_dgrad(_p(dy), _p(w), _p(dx), F, H, W, C, k, stride, _s()))


def dwconv_wgrad(dy, x, dw, F, H, W, C, k, stride):
    check(lib.lr_dwconv_wgrad(_p(dy), _p(x), _p(dw), F, H, W, C, k, stride, _s()))


def bn_act_fwd(x, stats, bn, act, training, z, rows, C, residual=None, res_pre=False):
    check(lib.lr_bn_act_fwd(_p(x), _p(stats), _p(bn.weight), _p(bn.bias), _p(bn.running_mean), _p(bn.running_var),
                            _p(bn.num_batches_tracked), bn.eps, bn.momentum, act, int(training), _p(residual),
                            int(res_pre), _p(z), rows, C, _s()))


def bn_act_bwd(x, stats, bn, act, training, dz, sums, dx, dgamma, dbeta, rows, C, z_out=None, dres=None):
    check(lib.lr_bn_act_bwd(_p(x), _p(stats), _p(bn.weight), _p(bn.bias), _p(bn.running_mean), _p(bn.running_var),
                            bn.eps, act, int(training), _p(dz), _p(z_out), _p(dres), _p(sums), _p(dx), _p(dgamma),
                            _p(dbeta), rows, C, _s()))


def nhwc_layout(F, H, W, C):
    """(is_u8, scale, F, T, sb, st, sc, sh, sw) of a channels-last float activation [F,H,W,C] for lr_im2col."""
    return (0, 1.0, F, 1, H * W * C, 0, 1, W * C, C)


def im2col(x, src, Hs, Ws, C, kh, kw, stride, pad, transposed, Hd, Wd, col, ldk):
    """src = (is_u8, scale, F, T, sb, st, sc, sh, sw)."""
    check(lib.lr_im2col(_p(x), *src, Hs, Ws, C, kh, kw, stride, pad, int(transposed), Hd, Wd, _p(col), ldk, _s()))


def weight_transpose(w, wt, Cout, Cin, kk, ldt):
    check(lib.lr_weight_transpose(_p(w), _p(wt), Cout, Cin, kk, ldt, _s()))


def maxpool_fwd(x, y, arg, F, H, W, C, k, stride, pad):
    check(lib.lr_maxpool_fwd(_p(x), _p(y), _p(arg), F, H, W, C, k, stride, pad, _s()))


def maxpool_bwd(dy, arg, dx, F, H, W, C, k, stride, pad):
    check(lib.lr_maxpool_bwd(_p(dy), _p(arg), _p(dx), F, H, W, C, k, stride, pad, _s()))


def dropout_fwd(x, y, mask, n, p, seed, step):
    check(lib.lr_dropout_fwd(_p(x), _p(y), _p(mask), n, p, seed, _p(step), _s()))


def dropout_bwd(dy, mask, dx, n, p):
    check(lib.lr_dropout_bwd(_p(dy), _p(mask), _p(dx), n, p, _s()))


def rng_tick(step):
    check(lib.lr_rng_tick(_p(step), _s()))


def frame_reduce(a, g, p, F, HW, C, mode):
    check(lib.lr_frame_reduce(_p(a), _p(g), _p(p), F, HW, C, mode, _s()))


def linear_small_fwd(x, ldx, w, b, y, ldy, M, N, K, act=ACT_NONE):
    """y = act(x w^T + b) on M <= 32 rows."""
    check(lib.lr_linear_small_fwd(_p(x), ldx, _p(w), _p(b), _p(y), ldy, M, N, K, act, _s()))


def linear_small_dgrad(dy, ldy, w, dx, ldx, M, N, K, r=None, ldr=0):
    """dx = dy w (+ r) on M <= 32 rows."""
    check(lib.lr_linear_small_dgrad(_p(dy), ldy, _p(w), _p(dx), ldx, _p(r), ldr, M, N, K, _s()))


def se_fc_fwd(p, w1, b1, w2, b2, h1, s, F, C, Cs, act1=ACT_RELU, act2=ACT_HSIGMOID):
    """h1 = act1(p w1^T + b1), s = act2(h1 w2^T + b2): the SE gate of the pooled vectors in one launch."""
    check(lib.lr_se_fc_fwd(_p(p), _p(w1), _p(b1), _p(w2), _p(b2), _p(h1), _p(s), F, C, Cs, act1, act2, _s()))


def se_fc_bwd(ds, s, h1, w1, w2, dz1, dp, F, C, Cs, act1=ACT_RELU, act2=ACT_HSIGMOID):
    """ds <- dz2 = ds * act2'(s); dz1 = (dz2 w2) * act1'(h1); dp = dz1 w1."""
    check(lib.lr_se_fc_bwd(_p(ds), _p(s), _p(h1), _p(w1), _p(w2), _p(dz1), _p(dp), F, C, Cs, act1, act2, _s()))


def frame_scale(a, s, dp, out, F, HW, C):
    check(lib.lr_frame_scale(_p(a), _p(s), _p(dp), _p(out), F, HW, C, _s()))


def act_bwd(dy, y, n, act):
    check(lib.lr_act_bwd(_p(dy), _p(y), n, act, _s()))


def colsum(dY, ld, M, N, db):
    check(lib.lr_colsum(_p(dY), ld, M, N, _p(db), _s()))


def lstm_fwd(xproj, ldx, bhh, whh, out, ldo, gates, cst, hprev, B, T, H, nsteps, reverse):
    check(lib.lr_lstm_fwd(_p(xproj), ldx, _p(bhh), _p(whh), _p(out), ldo, _p(gates), _p(cst), _p(hprev), B, T, H,
                          nsteps, int(reverse), _s()))


def lstm_bwd(dout, ldo, dout_step, gates, cst, whh, dgates, B, T, H, nsteps, reverse):
    check(lib.lr_lstm_bwd(_p(dout), ldo, dout_step, _p(gates), _p(cst), _p(whh), _p(dgates), B, T, H, nsteps,
                          int(reverse), _s()))


def lstm_fwd_tc(xproj, ldx, bhh, whh, out, ldo, gates, cst, hprev, B, T, H, nsteps, reverse):
    check(lib.lr_lstm_fwd_tc(_p(xproj), ldx, _p(bhh), _p(whh), _p(out), ldo, _p(gates), _p(cst), _p(hprev), B, T, H,
                             nsteps, int(reverse), _s()))


def lstm_bwd_tc(dout, ldo, dout_step, gates, cst, whh, dgates, B, T, H, nsteps, reverse):
    check(lib.lr_lstm_bwd_tc(_p(dout), ldo, dout_step, _p(gates), _p(cst), _p(whh), _p(dgates), B, T, H, nsteps,
                             int(reverse), _s()))


def audio_conv_fwd(x, w, bias, out, ldo, arg, B, H, W):
    check(lib.lr_audio_conv_fwd(_p(x), _p(w), _p(bias), _p(out), ldo, _p(arg), B, H, W, _s()))


def audio_conv_bwd(x, dA, lda, arg, dw, db, B, H, W):
    check(lib.lr_audio_conv_bwd(_p(x), _p(dA), lda, _p(arg), _p(dw), _p(db), B, H, W, _s()))


def ce_loss(logits, labels, loss, dlogits, correct, B, C, inv_n):
    check(lib.lr_ce_loss(_p(logits), _p(labels), _p(loss), _p(dlogits), _p(correct), B, C, inv_n, _s()))


def adam_step(p, g, m, v, state, n, beta1, beta2, eps, weight_decay, grad_scale):
    check(lib.lr_adam_step(_p(p), _p(g), _p(m), _p(v), _p(state), n, beta1, beta2, eps, weight_decay, grad_scale, _s()))


def copy2d(dst, ldd, src, lds, rows, cols):
    check(lib.lr_copy2d(_p(dst), ldd, _p(src), lds, rows, cols, _s()))


def gemm_tf32(A, lda, a_trans, B, ldb, b_trans, C, ldc, M, N, K, bias=None, act=ACT_NONE, R=None, ldr=0, stats=None,
              ksplit=1):
    """Tensor-core GEMM (TF32 products, fp32 accumulate), same layout flags / epilogue as gemm()."""
    check(lib.lr_gemm_tf32(_p(A), lda, a_trans, _p(B), ldb, b_trans, _p(C), ldc, M, N, K, _p(bias), act, _p(R), ldr,
                           _p(stats), ksplit, _s()))


def im2col_tap(x, F, Hs, Ws, C, kh, kw, stride, pad, transposed, Hd, Wd, col, pad_w=None):
    check(lib.lr_im2col_tap(_p(x), F, Hs, Ws, C, kh, kw, stride, pad, pad if pad_w is None else pad_w, int(transposed),
                            Hd, Wd, _p(col), _s()))


def weight_tap(src, dst, Cout, Cin, kk, mode):
    check(lib.lr_weight_tap(_p(src), _p(dst), Cout, Cin, kk, mode, _s()))
