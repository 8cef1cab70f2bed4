"""Adds node-correlation fixtures to tests/golden/ (run once in the build container, where
/root/reference exists): the published cambium2 node_correlations.csv and outputs of the
reference's own attractor_analysis.compute_frequency_spearmanrho on seeded random inputs."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refshim  # noqa: E402  (stubs ZODB & co. and puts /root/reference on sys.path)
_refshim.install()
from boolsi import attractor_analysis as ref  # noqa: E402

pub_path = os.path.join(HERE, 'published.json')
pub = json.load(open(pub_path))
with open('/root/reference/examples/output8_cambium2/node_correlations.csv') as f:
    pub['files']['output8_cambium2/node_correlations.csv'] = f.read()
json.dump(pub, open(pub_path, 'w'))

cases = []
rng = np.random.default_rng(2024)
for n_obs, n_nodes in [(5, 4), (12, 9), (39, 30), (3, 6)]:
    # cycle-mean activities: rationals with small denominators, many ties, some constant columns
    lam = rng.integers(1, 7, size=n_obs)
    data = np.array([[rng.integers(0, l + 1) / l for _ in range(n_nodes)] for l in lam])
    data[:, 0] = 0.0
    freq = rng.integers(1, 2000, size=n_obs)
    rho, p = ref.compute_frequency_spearmanrho(data, freq)
    cases.append({'data': data.tolist(), 'frequencies': freq.tolist(),
                  'rho': [[None if np.isnan(x) else float(x) for x in row] for row in rho],
                  'p': [[None if np.isnan(x) else float(x) for x in row] for row in p]})
json.dump({'generator': 'tests/golden/add_correlations.py: reference attractor_analysis.py:70-145',
           'cases': cases}, open(os.path.join(HERE, 'correlations.json'), 'w'))
print('ok', len(cases))
