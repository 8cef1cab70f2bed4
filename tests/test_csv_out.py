"""CPU: the CSV write-out (boolsi_b200/csv_out.py) against the reference's published example files,
fed with the reference-generated golden results (no GPU involved)."""
import csv
import io

import pytest

import codec
import goldenlib
from boolsi_b200 import ResultStore, csv_out, init_attractor_db_structure, init_simulation_db_structure
from boolsi_b200.attract import AggregatedAttractor
from boolsi_b200.problems import state_from_code
from boolsi_b200.simulate import Simulation, store_simulations

PUBLISHED = codec.load('published.json')['files']


@pytest.mark.parametrize('case_name,outdir', [
    ('simulate_example1', 'output2_example1'),
    ('simulate_example1_fixed_nodes', 'output5_example1_fixed_nodes'),
    ('simulate_example1_perturbations', 'output6_example1_perturbations'),
    ('simulate_cambium1', 'output7_cambium1')])
def test_simulation_csvs(case_name, outdir):
    case = [c for c in goldenlib.cases('simulate') if c['name'] == case_name][0]
    a = goldenlib.case_args(case)
    db = ResultStore()
    init_simulation_db_structure(db)
    store_simulations(db, [Simulation([state_from_code(c, a['n_nodes']) for c in codes], fixed, pert)
                           for _, codes, fixed, pert in goldenlib.expected_simulations(case)])
    summaries, states = csv_out.simulations_csv(db, goldenlib.networks()[case['network']][0])
    assert states.replace('\r\n', '\n') == PUBLISHED[outdir + '/simulations.csv']
    assert summaries.replace('\r\n', '\n') == PUBLISHED[outdir + '/simulation_summaries.csv']


def test_attractor_csvs_example2_values():
    """examples/output3_example2 is in an older summary format (ids `attractor_1`, rounded mean):
    states and the value columns must still agree."""
    case = [c for c in goldenlib.cases('attract') if c['name'] == 'attract_example2'][0]
    a = goldenlib.case_args(case)
    db = ResultStore()
    init_attractor_db_structure(db)
    for key, n, s1, s2, cyc in goldenlib.expected_attractor_rows(case):
        db.root.aggregated_attractors[(-n, key)] = AggregatedAttractor(
            key, [state_from_code(c, a['n_nodes']) for c in cyc], n, s1, s2)
        db.root.total_frequency.change(n)
    summaries, states = csv_out.attractors_csv(db, {}, ['A', 'B', 'C'], a['n_problems'])
    assert states.replace('\r\n', '\n') == PUBLISHED['output3_example2/attractors.csv']
    got = list(csv.reader(io.StringIO(summaries)))
    want = list(csv.reader(io.StringIO(PUBLISHED['output3_example2/attractor_summaries.csv'])))
    assert len(got) == len(want)
    for g, w in zip(got[1:], want[1:]):
        assert g[1] == w[1] and float(g[4]) == float(w[4])
        assert abs(float(g[2]) - float(w[2])) < 1e-9


def test_no_attractor_row():
    db = ResultStore()
    init_attractor_db_structure(db)
    db.root.aggregated_attractors[(-3, 5)] = AggregatedAttractor(5, [[True, False, True]], 3, 6, 14)
    db.root.total_frequency.set(3)
    summaries, _ = csv_out.attractors_csv(db, {1: False}, ['A', 'B', 'C'], 8, max_attractor_l=4, max_t=10)
    rows = list(csv.reader(io.StringIO(summaries)))
    assert rows[1] == ['no_attractor', '<= 4', '<= 10', '', '0.625']       # output.py:655-660
    assert rows[2][:2] == ['attractor1', '1'] and float(rows[2][4]) == 3 / 8


def test_vectorised_state_rows_equal_the_cell_by_cell_writer():
    """The byte-matrix formatter against csv.writer over the per-cell rows: random trajectories,
    fixed-node `_` markers, perturbation `*` markers (incl. on fixed nodes and at t = 0), both
    label styles, one node, empty trajectory."""
    import csv
    import io
    import numpy as np
    rng = np.random.default_rng(5)
    for trial in range(60):
        n_nodes = int(rng.integers(1, 40))
        n_rows = int(rng.integers(0, 30))
        states = rng.integers(0, 2, size=(n_rows, n_nodes)).astype(bool).tolist()
        fixed = {int(i): bool(rng.integers(0, 2)) for i in rng.choice(n_nodes, size=int(rng.integers(0, min(4, n_nodes) + 1)), replace=False)}
        perturbed = {}
        for _ in range(int(rng.integers(0, 5))):
            perturbed.setdefault(int(rng.integers(0, max(n_rows, 1) + 2)), {})[int(rng.integers(0, n_nodes))] = True
        labels = None if trial % 2 else ['t'] + ['t+%d' % t for t in range(1, n_rows)]
        want = io.StringIO(newline='')
        csv.writer(want).writerows(csv_out._state_rows(states, fixed, perturbed, 'simulation%d' % trial, labels))
        assert csv_out._state_rows_text(states, fixed, perturbed, 'simulation%d' % trial, labels) == want.getvalue()


def test_simulation_record_from_a_state_array_equals_the_one_from_lists():
    """Device results unpack to bool arrays: the record keeps the array for the CSV writer and
    hands the reference's list-of-lists out on demand; both routes give the same files."""
    import numpy as np
    from boolsi_b200 import _lib
    from boolsi_b200.simulate import Simulation
    rng = np.random.default_rng(9)
    words = rng.integers(0, 1 << 63, size=(12, 1), dtype=np.uint64)
    arr = _lib.state_array_from_word_rows(words, 40)
    assert arr.shape == (12, 40) and arr.dtype == bool
    assert arr.tolist() == _lib.states_from_word_rows(words, 40)
    assert [sum(int(b) << i for i, b in enumerate(row)) for row in arr.tolist()] == \
        [int(w) & ((1 << 40) - 1) for w in words[:, 0]]
    fixed, perturbed = {3: True}, {2: {5: False}, 7: {3: True, 0: True}}
    names = ['n%d' % i for i in range(40)]
    texts = []
    for states in (arr, arr.tolist()):
        db = ResultStore()
        init_simulation_db_structure(db)
        db.root.simulations[(0, 0)] = Simulation(states, fixed, perturbed)
        db.root.simulations[(0, 1)] = Simulation(states[:1], {}, {})
        assert db.root.simulations[(0, 0)].states == arr.tolist()
        texts.append(csv_out.simulations_csv(db, names))
    assert texts[0] == texts[1]
    assert texts[0][1].count('\r\n') == 1 + 12 + 1
    assert '1_*' in texts[0][1] or '0_*' in texts[0][1]
