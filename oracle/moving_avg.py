"""Oracle (test infrastructure): EMA normaliser / EMA estimate.

Restates /root/reference/src/madrona_learn/moving_avg.py:7-198 in NumPy float32.
State is a plain dict with the reference's keys: mu, inv_sigma, sigma, mu_biased,
sigma_sq_biased (float32 [dim]) and N (int32 scalar).
"""
import numpy as np

F32 = np.float32


def _f32(x):
    return np.asarray(x, F32)


class EMANormalizer:
    """ml/moving_avg.py:48-198."""

    def __init__(self, decay, eps=1e-5, disable=False):
        self.decay = decay
        self.eps = eps
        self.disable = disable

    def init_estimates(self, dim):
        # :56-75 -- mu 0, sigma 1 is a no-op normaliser until the first update.
        return dict(
            mu=np.zeros(dim, F32), inv_sigma=np.ones(dim, F32), sigma=np.ones(dim, F32),
            mu_biased=np.zeros(dim, F32), sigma_sq_biased=np.zeros(dim, F32),
            N=np.int32(0))

    def normalize(self, est, x):
        # :77-85
        x = _f32(x)
        return ((x - est['mu']) * est['inv_sigma']).astype(F32)

    def invert(self, est, x):
        # :87-95 (inv_dtype = float32 for the value normaliser, ml/train_state.py:307-316)
        x = _f32(x)
        return (x * est['sigma'] + est['mu']).astype(F32)

    def init_input_stats(self, est):
        # :97-101
        return np.zeros_like(est['mu']), np.zeros_like(est['mu'])

    def update_input_stats(self, cur_stats, num_prev_updates, x):
        """Equal-weight Chan merge of one more batch's (mean, var).  :103-129."""
        a_mean, a_var = cur_stats
        x = _f32(x)
        dim = x.shape[-1]
        flat = x.reshape(-1, dim)
        # reference: jnp.mean(..., dtype=f32) -- reduction order unspecified; oracle sums in
        # f64 and rounds once (tolerance rel 1e-5).
        b_mean = flat.mean(axis=0, dtype=np.float64).astype(F32)
        b_var = np.square((flat - b_mean).astype(F32)).mean(axis=0, dtype=np.float64).astype(F32)
        delta = (b_mean - a_mean).astype(F32)
        n_ab = num_prev_updates + 1
        b_weight = F32(1.0) / F32(n_ab)
        a_weight = F32(1) - b_weight
        ab_mean = (a_mean + delta * b_weight).astype(F32)
        ab_var = (a_weight * a_var + b_weight * b_var +
                  np.square(delta) * a_weight * b_weight).astype(F32)
        return ab_mean, ab_var

    def update_estimates(self, est, input_stats):
        """EMA of mean and variance with cross term and bias correction.  :131-181."""
        x_mean, x_var = input_stats
        mean_delta = (x_mean - est['mu']).astype(F32)
        one_minus_alpha = F32(self.decay)
        alpha = F32(1) - one_minus_alpha
        N = np.int32(est['N'])
        new_N = np.int32(N + 1)
        new_mu_biased = (one_minus_alpha * est['mu_biased'] + alpha * x_mean).astype(F32)
        new_sigma_sq_biased = (
            one_minus_alpha * est['sigma_sq_biased'] + alpha * x_var +
            (F32(N) / F32(new_N)) * (one_minus_alpha * alpha) * np.square(mean_delta)
        ).astype(F32)
        bias_correction = F32(-1) / np.expm1(F32(new_N) * np.log(one_minus_alpha), dtype=F32)
        new_mu = (new_mu_biased * bias_correction).astype(F32)
        new_sigma_sq = (new_sigma_sq_biased * bias_correction).astype(F32)
        new_inv_sigma = (F32(1) / np.sqrt(np.maximum(new_sigma_sq, F32(self.eps)), dtype=F32)).astype(F32)
        new_sigma = (F32(1) / new_inv_sigma).astype(F32)
        return dict(mu=new_mu, inv_sigma=new_inv_sigma, sigma=new_sigma,
                    mu_biased=new_mu_biased, sigma_sq_biased=new_sigma_sq_biased, N=new_N)

    def normalize_and_update_estimates(self, est, inputs):
        # :183-192 -- stats of this batch alone, EMA update, normalise with the NEW stats.
        stats = self.update_input_stats(self.init_input_stats(est), 0, inputs)
        est = self.update_estimates(est, stats)
        return est, self.normalize(est, inputs)


class EMAEstimate:
    """Scalar EMA with bias correction.  ml/moving_avg.py:7-44."""

    def __init__(self, decay, eps=1e-5):
        self.decay = decay
        self.eps = eps

    def init_estimates(self, dim=1):
        return dict(mu=np.zeros(dim, F32), mu_biased=np.zeros(dim, F32), N=np.int32(0))

    def update_estimates(self, est, x):
        x_mean = F32(np.mean(np.asarray(x, F32), dtype=np.float64))
        one_minus_alpha = F32(self.decay)
        alpha = F32(1) - one_minus_alpha
        new_N = np.int32(est['N'] + 1)
        new_mu_biased = (one_minus_alpha * est['mu_biased'] + alpha * x_mean).astype(F32)
        bias_correction = F32(-1) / np.expm1(F32(new_N) * np.log(one_minus_alpha), dtype=F32)
        return dict(mu=(new_mu_biased * bias_correction).astype(F32),
                    mu_biased=new_mu_biased, N=new_N)
