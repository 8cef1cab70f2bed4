"""GPU: the reference's PUBLISHED example output for `attract examples/cambium2.yaml`
(2^30 initial states, 39 attractors, examples/output8_cambium2/*.csv, carried as text in
tests/golden/published.json) against a full run of the CUDA path."""
import csv
import io
import math
from math import inf

import pytest

import codec
import boolsi_b200
from boolsi_b200.constants import NodeStateRange

pytestmark = pytest.mark.gpu


class _Comm:
    n_workers = 0
    rank = 0


def test_cambium2_full_run_matches_published_csv():
    pub = codec.load('published.json')
    _, preds, tts = codec.decode_network(codec.load('networks.json')['networks']['cambium'])
    origin, variations = codec.decode_problem(pub['cambium2_problem'], NodeStateRange)
    n = pub['cambium2_n_problems']
    assert n == 2 ** 30
    db = boolsi_b200.ResultStore()
    found = boolsi_b200.attract_master(_Comm(), 100, origin, variations, preds, tts, inf, inf, n, True,
                                       db, False, None)
    summaries = list(csv.DictReader(io.StringIO(pub['files']['output8_cambium2/attractor_summaries.csv'])))
    rows = list(csv.reader(io.StringIO(pub['files']['output8_cambium2/attractors.csv'])))[1:]
    assert found == len(summaries) == 39
    assert db.root.total_frequency() == n                      # every initial state reaches an attractor
    got = list(db.root.aggregated_attractors.values())
    published_states = {}
    for r in rows:
        published_states.setdefault(r[0], []).append([c.rstrip('_*') == '1' for c in r[2:]])
    for i, (att, s) in enumerate(zip(got, summaries)):
        assert s['attractor_id'] == 'attractor%d' % (i + 1)
        assert len(att.states) == int(s['length'])
        assert att.frequency / n == float(s['relative_frequency'])          # exact
        # the reference accumulates mean / variance in floats (order-dependent noise ~1e-14)
        assert math.isclose(att.trajectory_l_mean, float(s['trajectory_length_mean']), rel_tol=1e-12)
        sd = math.sqrt(att.trajectory_l_variation_sum / (att.frequency - 1))   # output.py:668-670
        assert math.isclose(sd, float(s['trajectory_length_SD']), rel_tol=1e-9)
        assert att.states == published_states[s['attractor_id']]
