"""Attractor node-activity and node-correlation inputs.

What the reference computes in `attractor_analysis.py:30-145` from `db_conn.root.aggregated_attractors`:
each attractor becomes one observation in [0,1]^N (mean activity of every node over the cycle,
`attractor_analysis.py:49-51`), weighted by its frequency; node correlations are frequency-weighted
Spearman rho with two-sided t-test p-values.  The reference module itself works unchanged on
`boolsi_b200.ResultStore` (it only reads `.states` and `.frequency`); this restatement exists so
that the GPU path can hand the observation matrix over without building Python lists of lists, and
so that the C3 configuration ("attract cambium2 with node-correlation analysis") is checkable on
the GPU box where the reference package is absent.
"""
import csv
import io

import numpy as np
from scipy import stats


def node_activity(attractors):
    """-> (observations [n_attractors, N] float64, frequencies [n_attractors] int64)
    (attractor_analysis.py:45-54).  Attractors that come off the GPU path carry `node_on_counts` -- per
    node, the number of cycle states in which it is on, counted on the device -- and the observation
    matrix is ONE array division (counts / cycle length: the same float64 values as the reference's
    `np.mean(states, axis=0)`, a sum of 0/1 divided by the length); anything else falls back to the
    states."""
    frequencies = np.array([a.frequency for a in attractors], dtype=np.int64)
    counts = [getattr(a, 'node_on_counts', None) for a in attractors]
    if attractors and all(c is not None for c in counts):
        lengths = np.array([a.n_states if hasattr(a, 'n_states') else len(a.states) for a in attractors],
                           dtype=np.float64)
        observations = np.asarray(counts, dtype=np.float64) / lengths[:, None]
    else:
        observations = np.array([np.asarray(a.states, dtype=np.float64).mean(axis=0) for a in attractors])
    return observations, frequencies


def frequency_ranks(values, frequencies):
    """Average ranks of `values` when observation i is repeated frequencies[i] times
    (attractor_analysis.py:85-114)."""
    order = np.argsort(values, kind='stable')
    v, f = values[order], frequencies[order]
    new_group = np.concatenate(([True], v[1:] != v[:-1]))
    group = np.cumsum(new_group) - 1
    group_freq = np.bincount(group, weights=f).astype(np.int64)
    before = np.concatenate(([0], np.cumsum(group_freq)[:-1]))
    group_rank = before + 0.5 * (group_freq + 1)
    ranks = np.empty(len(values), dtype=np.float64)
    ranks[order] = group_rank[group]
    return ranks


def frequency_spearman(observations, frequencies):
    """-> (Rho [N, N], P [N, N]) (attractor_analysis.py:70-82, 117-145)."""
    ranks = np.column_stack([frequency_ranks(observations[:, j], frequencies)
                             for j in range(observations.shape[1])])
    df = frequencies.sum() - 2
    sigma = np.cov(ranks.T, fweights=frequencies)
    diag = sigma.diagonal()
    with np.errstate(invalid='ignore', divide='ignore'):
        rho = sigma / np.sqrt(np.multiply.outer(diag, diag))
        t = rho / np.sqrt((1 - rho * rho) / df)
        p = 2 * stats.t.sf(np.abs(t), df)
    return rho, p


def find_node_correlations(db_conn):
    """Drop-in for `attractor_analysis.find_node_correlations` (without the progress bar):
    None when there is a single attractor or total frequency <= 2 (attractor_analysis.py:8-27)."""
    root = db_conn.root
    if root.n_aggregated_attractors() == 1 or root.total_frequency() <= 2:
        return None
    observations, frequencies = node_activity(list(root.aggregated_attractors.values()))
    return frequency_spearman(observations, frequencies)


def node_correlations_csv(rho, p, p_value, node_names):
    """`node_correlations.csv` text: significant, nonsignificant, absent correlations, in this
    order and with the reference's sort keys (output.py:1072-1087)."""
    out = io.StringIO(newline='')
    writer = csv.writer(out)
    writer.writerow(['node_1', 'node_2', 'rho', 'p_value'])
    na = np.isnan(rho)
    with np.errstate(invalid='ignore'):
        significant = p < p_value
    nonsignificant = np.logical_and(~na, ~significant)
    keys = [(significant, lambda c: (1 - abs(c[1]), c[0][0], c[0][1])),
            (nonsignificant, lambda c: (c[2], 1 - abs(c[1]), c[0][0], c[0][1])),
            (na, lambda c: (c[0][0], c[0][1]))]
    for mask, key in keys:
        triu = np.triu(mask, 1)
        for (n1, n2), r, pv in sorted(zip(np.argwhere(triu), rho[triu], p[triu]), key=key):
            writer.writerow([node_names[n1], node_names[n2], r, pv])
    return out.getvalue()
