"""ESS / MCSE for many independent chains (the reference has no ESS code, SURVEY.md section 5;
the headline metric asks for ESS/sec, so the estimator is stated here).

With B independent chains of S draws each, per coordinate
    m_c   = chain means,            W = mean within-chain variance,
    Vb    = var_c(m_c)  (ddof=1)    -> Var of one chain's mean, estimated ACROSS chains,
    var+  = W (S-1)/S + Vb          (pooled posterior variance, Gelman et al.),
    ESS   = B * var+ / Vb           (capped at B*S),     MCSE(mean) = sqrt(Vb / B).
Because the chains are independent, Vb is an unbiased estimate of the Monte Carlo variance of
a chain mean without any autocorrelation model (batch means with batch = chain).
"""
from __future__ import annotations

import torch


def chain_summary(s1, s2, S):
    """From per-chain sums (B, D) over S draws: dict of pooled mean, var, mcse, ess (all (D,))."""
    B = s1.shape[0]
    s1, s2 = s1.double(), s2.double()
    m = s1 / S
    within = (s2 - S * m * m) / max(S - 1, 1)
    W = within.mean(0)
    mean = m.mean(0)
    Vb = m.var(0, unbiased=True) if B > 1 else torch.full_like(mean, float("nan"))
    var_plus = W * (S - 1) / S + Vb
    ess = torch.clamp(B * var_plus / Vb, max=float(B * S))
    # MCSE of the pooled variance estimate from the spread of per-chain variances
    mcse_var = torch.sqrt(within.var(0, unbiased=True) / B) if B > 1 else torch.full_like(mean, float("nan"))
    return dict(mean=mean, var=var_plus, mcse_mean=torch.sqrt(Vb / B), mcse_var=mcse_var, ess=ess,
                ess_per_draw=ess / (B * S), rhat=torch.sqrt(var_plus / W))
