"""TEST INFRASTRUCTURE (parity oracle, CPU): restatement of the seeded k-means that replaces `stats::kmeans(centers = k,
nstart)` inside ini_kmeans (MAGMAclust.R:562-598). Only tests/ may import this module.

"Parity unpinned" against R: stats::kmeans is Hartigan-Wong on R's global RNG and the reference holds no stored partition,
so no artefact constrains it. What this oracle pins, bit for bit, is the CUDA implementation (magmaclust_b200/csrc/kmeans.cu):
the same prescribed order of every floating-point operation --
  distance  d2(i, k): terms (x_id - c_kd)^2 added for d = 0, 1, ... in that order
  label     first k of the smallest d2
  centre    per (k, d): rows in chunks of CHUNK, each chunk summed in row order, chunk sums added in chunk order, / count;
            an empty cluster keeps its centre
  cost      sum_i d2(i, label_i), same two-level order
  Lloyd     assign; stop if no label moved; update; at most n_iter assignment passes; best start = first of the smallest cost
"""
import numpy as np

CHUNK = 1024


def _d2(feat, cent):
    """[M][K] squared distances, terms added in dimension order (np.sum's pairwise blocks would reorder them)"""
    out = np.zeros((feat.shape[0], cent.shape[0]))
    for d in range(feat.shape[1]):
        df = feat[:, d][:, None] - cent[:, d][None, :]
        out = out + df * df
    return out


def _seq_sum(v):
    """left-to-right sum (np.cumsum accumulates sequentially); adding the +0.0 of masked-out rows changes nothing"""
    return np.cumsum(v, axis=0)[-1] if v.shape[0] else np.zeros(v.shape[1:])


def _two_level_sum(v):
    parts = np.array([_seq_sum(v[s:s + CHUNK]) for s in range(0, v.shape[0], CHUNK)])
    return _seq_sum(parts)


def lloyd(feat, cent, n_iter=100):
    feat = np.asarray(feat, dtype=np.float64)
    cent = np.array(cent, dtype=np.float64)
    lab = np.full(feat.shape[0], -1)
    iters = 0
    for _ in range(n_iter):
        new_lab = np.argmin(_d2(feat, cent), axis=1)
        iters += 1
        if np.array_equal(new_lab, lab):
            break
        lab = new_lab
        for k in range(cent.shape[0]):
            m = lab == k
            if m.any():
                cent[k] = _two_level_sum(np.where(m[:, None], feat, 0.0)) / float(m.sum())
    d2 = _d2(feat, cent)[np.arange(feat.shape[0]), lab]
    return lab, cent, float(_two_level_sum(d2[:, None])[0]), iters


def kmeans(feat, K, start_rows, n_iter=100):
    """(labels, centres, cost, max passes over the starts) of the best of the starts given by start_rows [nstart][K]"""
    feat = np.asarray(feat, dtype=np.float64)
    best, iters = None, 0
    for rows in np.asarray(start_rows):
        lab, cent, cost, it = lloyd(feat, feat[rows], n_iter)
        iters = max(iters, it)
        if best is None or cost < best[2]:
            best = (lab, cent, cost)
    return best[0], best[1], best[2], iters
