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
