"""Adds `-r` (reference-point detection, reference attract.py:305-371 / model.py:216-228) fixtures
to tests/golden/ (run once in the build container, where /root/reference exists).

For every attract case of cases.json the UNMODIFIED reference's `attract_master` is run with
`storing_all_states=False` (what `boolsi attract -r` passes, cli.py:256-261).  Recorded per case:
the attractors it reports and whether they equal the default-mode fixture.  The two modes agree
whenever every repeat is detected by a reference point by `max_t` (always so without a time cap);
they differ at the cap, where `-r` needs the repeat to be DETECTED by max_t and otherwise builds a
pseudo-attractor from `t - last_reference_t` without checking `attractor_is_found`
(SURVEY.md section 7.2).  boolsi_b200 follows the default mode and says so in a warning."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(REPO, 'tests'))
sys.path.insert(0, REPO)
import _refshim  # noqa: E402
_refshim.install()
from boolsi.attract import attract_master, init_attractor_db_structure  # noqa: E402
import codec  # noqa: E402
import goldenlib  # noqa: E402

out = []
for case in goldenlib.cases('attract'):
    a = goldenlib.case_args(case)
    db = _refshim.new_db_conn()
    init_attractor_db_structure(db)
    attract_master(_refshim.NoMpiComm(), 3, a['origin'], a['variations'], a['preds'], a['tts'], a['max_t'],
                   a['max_attractor_l'], a['n_problems'], False, db, False, None)
    rows = []
    for (neg_freq, key), obj in db.root.aggregated_attractors.items():
        codes = [codec.state_to_code(s) for s in obj.states]
        rows.append({'key': format(key, 'x'), 'states': [format(c, 'x') for c in codes],
                     'frequency': obj.frequency, 'mean': obj.trajectory_l_mean,
                     'variation_sum': obj.trajectory_l_variation_sum})
    default = [{'key': r['key'], 'states': r['states'], 'frequency': r['frequency']}
               for r in case['expected']['attractors']]
    same = default == [{'key': r['key'], 'states': r['states'], 'frequency': r['frequency']} for r in rows]
    if same:   # the statistics too (floats: the reference accumulates them in problem order in both modes)
        for r, d in zip(rows, case['expected']['attractors']):
            same = same and abs(r['mean'] - d['mean']) <= 1e-9 * max(1.0, abs(d['mean'])) and \
                abs(r['variation_sum'] - d['variation_sum']) <= 1e-6 * max(1.0, abs(d['variation_sum']))
    out.append({'name': case['name'], 'max_t': case['max_t'], 'equal_to_default_mode': same,
                'total_frequency': db.root.total_frequency(), 'attractors': rows})
    print('%-50s max_t=%-6s equal=%s' % (case['name'], case['max_t'], same), flush=True)
json.dump({'generator': 'tests/golden/add_rmode.py: reference attract_master(storing_all_states=False)',
           'cases': out}, open(os.path.join(HERE, 'rmode.json'), 'w'), indent=0)
print('ok', len(out), 'cases,', sum(c['equal_to_default_mode'] for c in out), 'equal to the default mode')
