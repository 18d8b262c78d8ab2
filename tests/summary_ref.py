"""Test infrastructure: a numpy stand-in for the two device reductions behind bayes_js_b200.summary.CudaBlockReducer
(amwg_summary_moments / amwg_summary_digit_hist), so that the host logic and the collectives run on CPU tensors.
Never imported by the product."""
import numpy as np


class NumpyBlockReducer:
    def moments(self, block):
        x = block.numpy()                                    # [rows, entries, chains]
        rows, entries, chains = x.shape
        m = x.sum(axis=0) / rows                             # [entries, chains]
        m2 = ((x - m[None]) ** 2).sum(axis=0)
        out = np.empty((entries, 4))
        for e in range(entries):
            mean = m[e].mean()
            out[e] = (chains, mean, ((m[e] - mean) ** 2).sum(), m2[e].sum())
        return out

    def digit_counts(self, block, npass, prefix_table):
        import torch
        from bayes_js_b200.summary import double_to_key
        x = block.numpy()
        rows, entries, chains = x.shape
        n_prefix = prefix_table.shape[1]
        counts = np.zeros((entries, n_prefix, 256), dtype=np.int64)
        shift = np.uint64(56 - 8 * npass)
        for e in range(entries):
            k = double_to_key(x[:, e, :].ravel())
            digit = ((k >> shift) & np.uint64(255)).astype(np.int64)
            hi = (k >> (shift + np.uint64(8))) if npass else np.zeros_like(k)
            for q in range(n_prefix):
                sel = digit if npass == 0 else digit[hi == prefix_table[e, q]]
                counts[e, q] = np.bincount(sel, minlength=256)
        return torch.from_numpy(counts)


def numpy_summary(x, probs):
    """x [rows, entries, chains] -> (mean, sd, rhat, quantiles) straight from numpy, as a user of sample() would compute them."""
    rows, entries, chains = x.shape
    flat = np.moveaxis(x, 1, 0).reshape(entries, -1)
    mean = flat.mean(axis=1)
    sd = flat.std(axis=1, ddof=1)
    W = x.var(axis=0, ddof=1).mean(axis=1) if rows > 1 else np.full(entries, np.nan)
    B_over_n = x.mean(axis=0).var(axis=1, ddof=1) if chains > 1 else np.full(entries, np.nan)
    with np.errstate(invalid="ignore", divide="ignore"):
        rhat = np.sqrt(((rows - 1) / rows * W + B_over_n) / W)
    q = np.quantile(flat, probs, axis=1) if len(probs) else np.empty((0, entries))
    return mean, sd, rhat, q
