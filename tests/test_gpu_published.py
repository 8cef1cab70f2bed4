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


# ---- byte-level parity of the CSV write-out with the reference's published example outputs --------
import goldenlib  # noqa: E402
from boolsi_b200 import csv_out  # noqa: E402

_PUBLISHED_SIMULATE = [('simulate_example1', 'output2_example1'),
                       ('simulate_example1_fixed_nodes', 'output5_example1_fixed_nodes'),
                       ('simulate_example1_perturbations', 'output6_example1_perturbations'),
                       ('simulate_cambium1', 'output7_cambium1')]


def _published(name):
    # the fixtures were read in text mode: CRLF rows arrive as LF
    return codec.load('published.json')['files'][name]


@pytest.mark.parametrize('case_name,outdir', _PUBLISHED_SIMULATE)
def test_simulations_csv_is_byte_identical_to_the_published_example(case_name, outdir):
    case = [c for c in goldenlib.cases('simulate') if c['name'] == case_name][0]
    a = goldenlib.case_args(case)
    node_names = goldenlib.networks()[case['network']][0]
    db = boolsi_b200.ResultStore()
    boolsi_b200.simulate_master(_Comm(), 100, a['origin'], a['variations'], a['preds'], a['tts'],
                                a['max_t'], a['n_problems'], db, None)
    summaries, states = csv_out.simulations_csv(db, node_names)
    assert states.replace('\r\n', '\n') == _published(outdir + '/simulations.csv')
    assert summaries.replace('\r\n', '\n') == _published(outdir + '/simulation_summaries.csv')
    assert states.count('\r\n') == states.count('\n')           # written with CRLF like the reference


def test_target_csv_is_byte_identical_to_the_published_example():
    case = [c for c in goldenlib.cases('target') if c['name'] == 'target_example3'][0]
    a = goldenlib.case_args(case)
    db = boolsi_b200.ResultStore()
    boolsi_b200.target_master(_Comm(), 100, a['origin'], a['variations'], a['target_code'],
                              a['target_nodes'], a['preds'], a['tts'], a['n_to_find'], a['max_t'],
                              a['n_problems'], db, None)
    _, states = csv_out.simulations_csv(db, goldenlib.networks()[case['network']][0])
    assert states.replace('\r\n', '\n') == _published('output4_example3/simulations.csv')


def test_cambium2_attractors_csv_is_byte_identical_to_the_published_example(tmp_path):
    pub = codec.load('published.json')
    node_names, preds, tts = codec.decode_network(codec.load('networks.json')['networks']['cambium'])
    origin, variations = codec.decode_problem(pub['cambium2_problem'], NodeStateRange)
    n = pub['cambium2_n_problems']
    db = boolsi_b200.ResultStore()
    boolsi_b200.attract_master(_Comm(), 100, origin, variations, preds, tts, inf, inf, n, True, db,
                               False, None)
    csv_out.write_attractor_csvs(db, origin[1], node_names, n, str(tmp_path))
    with open(tmp_path / 'attractors.csv', newline='') as f:
        states = f.read()
    assert states.replace('\r\n', '\n') == _published('output8_cambium2/attractors.csv')
    # summaries: every field identical except the two float statistics, which the reference
    # accumulates in floating point (order-dependent last digits)
    with open(tmp_path / 'attractor_summaries.csv', newline='') as f:
        got = list(csv.reader(f))
    want = list(csv.reader(io.StringIO(_published('output8_cambium2/attractor_summaries.csv'))))
    assert len(got) == len(want) and got[0] == want[0]
    for g, w in zip(got[1:], want[1:]):
        assert (g[0], g[1], g[4]) == (w[0], w[1], w[4])
        assert math.isclose(float(g[2]), float(w[2]), rel_tol=1e-12)
        assert math.isclose(float(g[3]), float(w[3]), rel_tol=1e-9)


def test_cambium2_node_correlations_match_the_published_csv():
    """BASELINE.json configs[2]: attract cambium2 over all 2^30 initial states WITH the attractor
    node-correlation analysis (reference cli.py:263-276 -> attractor_analysis.py, output.py:868)."""
    from boolsi_b200 import analysis
    pub = codec.load('published.json')
    node_names, preds, tts = codec.decode_network(codec.load('networks.json')['networks']['cambium'])
    origin, variations = codec.decode_problem(pub['cambium2_problem'], NodeStateRange)
    n = pub['cambium2_n_problems']
    db = boolsi_b200.ResultStore()
    boolsi_b200.attract_master(_Comm(), 100, origin, variations, preds, tts, inf, inf, n, True, db,
                               False, None)
    rho, p = analysis.find_node_correlations(db)
    text = analysis.node_correlations_csv(rho, p, 0.05, node_names)       # default --p-value 0.05
    got = list(csv.reader(io.StringIO(text)))
    want = list(csv.reader(io.StringIO(_published('output8_cambium2/node_correlations.csv'))))
    assert got[0] == want[0] and len(got) == len(want)
    by_pair = {(r[0], r[1]): r for r in got[1:]}
    for w in want[1:]:
        g = by_pair[(w[0], w[1])]
        for gv, wv in zip(g[2:], w[2:]):
            if wv == 'nan':
                assert gv == 'nan'
            else:
                assert math.isclose(float(gv), float(wv), rel_tol=1e-9, abs_tol=1e-300)
