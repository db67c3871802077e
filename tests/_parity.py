"""Shared parity measures for the bf16 (tensor-core) decoder path against the fp32 oracle.

What a bf16-operand GEMM chain can and cannot promise (tools/bf16_error_study.py): every GEMM
operand is rounded to 8 significant bits, so every logit carries a relative error of a few 1e-3 of the logit
spread; an arg-max decided by less than that flips.  On the trained-like synthetic captioner (logit std ~3,
median top-1/top-2 gap ~0.9 nats, hundreds of distinct words, the consumed word feeding back) about 1 % of the
decisions are that close, and in FREE-RUNNING greedy decoding a flipped word changes everything after it.  The
bars therefore are:
  * per decision (both models on the SAME prefix): agreement >= 98.5 %, and every disagreement is a near-tie of
    the fp32 oracle itself (its own gap between the two tokens below NEAR_TIE);
  * free-running: captions are identical up to a first divergence which is such a near-tie; the raw token
    agreement is reported and floored well above chance;
  * |delta log p| where the fp32 model has mass: rms <= 2e-2 (the north-star figure), tails bounded.
"""
import numpy as np

NEAR_TIE = 0.25          # nats; measured worst case 0.11 (v1), see tools/bf16_error_study.py


def log_softmax(z):
    m = z.max(-1, keepdims=True)
    return z - (np.log(np.exp(z - m).sum(-1, keepdims=True)) + m)


def first_divergence(tok, tok_want):
    """index of the first differing position per row (P if none)."""
    agree = tok == tok_want
    return np.where(agree.all(1), tok.shape[1], (~agree).argmax(1))


def divergence_gaps(tok, tok_want, z_want):
    """The oracle's own logit gap (its token minus the other path's token) at every row's first divergence."""
    first = first_divergence(tok, tok_want)
    rows = np.nonzero(first < tok.shape[1])[0]
    t = first[rows]
    return z_want[rows, t, tok_want[rows, t]] - z_want[rows, t, tok[rows, t]]


def decision_gaps(choice, tok_want, z_want):
    """Same-prefix decisions: the oracle's gap wherever the other path picked a different token."""
    r, t = np.nonzero(choice != tok_want)
    return z_want[r, t, tok_want[r, t]] - z_want[r, t, choice[r, t]]


def logp_errors(lp, lp_want, floor=-12.0):
    e = np.abs(lp - lp_want)[lp_want > floor]
    return dict(rms=float(np.sqrt((e.astype(np.float64) ** 2).mean())), p999=float(np.quantile(e, 0.999)), max=float(e.max()))
