"""CSV write-out of results in the reference's exact format.

Restates only the CSV half of reference `output.py`: the two files per mode that
`output_simulation_results` (output.py:307-457) writes -- `<name>_summaries.csv` and `<name>s.csv`
-- with `format_state` markers (`_` fixed node, `*` perturbation, output.py:1092-1110),
`write_states` rows (output.py:845-865), the summary rows of output.py:514-528 (simulations) and
output.py:667-680 (attractors) and the "no attractor" preamble row (output.py:655-660).  Files are
written as the reference does (`newline=''`, csv.writer default dialect, i.e. CRLF rows).
PDF / image output stays the reference's (matplotlib is not part of this path).
"""
import csv
import io
import math
import os

import numpy as np

simulation_name = 'simulation'              # constants.py:45
aggregated_attractor_name = 'attractor'     # constants.py:46


def _state_rows(states, fixed_nodes, perturbed_nodes_by_t, states_id, time_labels, rows=None):
    """write_states + format_state (output.py:845-865, 1092-1110), cell by cell."""
    fixed = list(fixed_nodes)
    out = []
    for t in (range(len(states)) if rows is None else rows):
        cells = ['1' if v else '0' for v in states[t]]
        for node in fixed:
            cells[node] += '_'
        for node in perturbed_nodes_by_t.get(t, ()):
            cells[node] += '*'
        out.append([states_id, time_labels[t] if time_labels else str(t)] + cells)
    return out


def _state_rows_text(states, fixed_nodes, perturbed_nodes_by_t, states_id, time_labels):
    """The same rows as CSV text, without a Python operation per cell: all rows of one trajectory
    share the positions of commas and `_` markers, so they are one byte matrix (a template row
    repeated, the 0/1 digits written into it column-wise); only rows with a perturbation marker
    (`*` changes the row's width) go cell by cell.  Ids, time labels and cells never contain a
    character that csv.writer would quote, so the text equals what csv.writer produces."""
    n_rows = len(states)
    if n_rows == 0:
        return ''
    arr = np.asarray(states, dtype=bool)
    if arr.ndim != 2 or arr.shape[1] == 0:
        buf = io.StringIO(newline='')
        csv.writer(buf).writerows(_state_rows(arr.tolist(), fixed_nodes, perturbed_nodes_by_t, states_id, time_labels))
        return buf.getvalue()
    n_nodes = arr.shape[1]
    fixed = set(fixed_nodes)
    template, digit_at = bytearray(), []
    for node in range(n_nodes):
        digit_at.append(len(template))
        template += b'0_' if node in fixed else b'0'
        template += b',' if node < n_nodes - 1 else b'\r\n'
    matrix = np.tile(np.frombuffer(bytes(template), np.uint8), (n_rows, 1))
    matrix[:, digit_at] = arr.astype(np.uint8) + ord('0')
    cells = matrix.view('S%d' % len(template)).ravel().tolist()
    prefix = states_id.encode('ascii') + b','
    labels = [label.encode('ascii') + b',' for label in time_labels] if time_labels else _plain_labels(n_rows)
    lines = [prefix + label + row for label, row in zip(labels, cells)]
    for t in perturbed_nodes_by_t:
        if 0 <= t < n_rows and perturbed_nodes_by_t[t]:
            row = _state_rows(arr[t:t + 1].tolist(), fixed_nodes, {0: perturbed_nodes_by_t[t]}, states_id,
                              [time_labels[t] if time_labels else str(t)])[0]
            lines[t] = ','.join(row).encode('ascii') + b'\r\n'
    return b''.join(lines).decode('ascii')


_LABELS = []


def _plain_labels(n_rows):
    """[b'0,', b'1,', ...] (cached: every simulation of a run has the same time axis)"""
    while len(_LABELS) < n_rows:
        _LABELS.append(b'%d,' % len(_LABELS))
    return _LABELS[:n_rows]


def simulations_csv(db_conn, node_names):
    """-> (simulation_summaries.csv text, simulations.csv text) for `db_conn.root.simulations`."""
    n_nodes = len(node_names)
    summaries, states = io.StringIO(newline=''), io.StringIO(newline='')
    ws, wt = csv.writer(summaries), csv.writer(states)
    ws.writerow([simulation_name + '_id', 'length'] +
                ['{}_is_fixed'.format(n) for n in node_names] +
                ['n_{}_perturbations'.format(n) for n in node_names])
    wt.writerow(['{}_id'.format(simulation_name), 'time'] + list(node_names))
    for index, simulation in enumerate(db_conn.root.simulations.values()):
        n_perturbations_by_node = [0] * n_nodes
        for perturbed_nodes in simulation.perturbed_nodes_by_t.values():
            for node in perturbed_nodes:
                n_perturbations_by_node[node] += 1
        sim_id = '{}{}'.format(simulation_name, index + 1)
        # (device results carry their states as a bool array; reference records as lists)
        sim_states = getattr(simulation, 'states_array', None)
        if sim_states is None:
            sim_states = simulation.states
        ws.writerow([sim_id, len(sim_states) - 1] +
                    [int(node in simulation.fixed_nodes) for node in range(n_nodes)] +
                    n_perturbations_by_node)
        states.write(_state_rows_text(sim_states, simulation.fixed_nodes,
                                      simulation.perturbed_nodes_by_t, sim_id, None))
    return summaries.getvalue(), states.getvalue()


def attractors_csv(db_conn, fixed_nodes, node_names, n_simulation_problems, max_attractor_l=None,
                   max_t=None):
    """-> (attractor_summaries.csv text, attractors.csv text) for `db_conn.root.aggregated_attractors`.
    `max_attractor_l` / `max_t` are None when unlimited, as the CLI passes them (cli.py:281-283)."""
    summaries, states = io.StringIO(newline=''), io.StringIO(newline='')
    ws, wt = csv.writer(summaries), csv.writer(states)
    ws.writerow([aggregated_attractor_name + '_id', 'length', 'trajectory_length_mean',
                 'trajectory_length_SD', 'relative_frequency'])
    wt.writerow(['{}_id'.format(aggregated_attractor_name), 'time'] + list(node_names))
    if db_conn.root.total_frequency() < n_simulation_problems:
        ws.writerow(['no_' + aggregated_attractor_name,
                     '' if max_attractor_l is None else '<= {}'.format(max_attractor_l),
                     '' if max_t is None else '<= {}'.format(max_t), '',
                     1 - db_conn.root.total_frequency() / n_simulation_problems])
    for index, attractor in enumerate(db_conn.root.aggregated_attractors.values()):
        if attractor.frequency > 1:
            sd = math.sqrt(attractor.trajectory_l_variation_sum / (attractor.frequency - 1))
        else:
            sd = math.nan
        att_id = '{}{}'.format(aggregated_attractor_name, index + 1)
        ws.writerow([att_id, len(attractor.states), attractor.trajectory_l_mean, sd,
                     attractor.frequency / n_simulation_problems])
        labels = ['t'] + ['t+{}'.format(t) for t in range(1, len(attractor.states))]
        states.write(_state_rows_text(attractor.states, fixed_nodes, {}, att_id, labels))
    return summaries.getvalue(), states.getvalue()


def _write(path, text):
    with open(path, 'w', newline='', encoding='utf-8') as f:
        f.write(text)


def write_simulation_csvs(db_conn, node_names, output_dirpath):
    os.makedirs(output_dirpath, exist_ok=True)
    summaries, states = simulations_csv(db_conn, node_names)
    _write(os.path.join(output_dirpath, simulation_name + '_summaries.csv'), summaries)
    _write(os.path.join(output_dirpath, simulation_name + 's.csv'), states)


def write_attractor_csvs(db_conn, fixed_nodes, node_names, n_simulation_problems, output_dirpath,
                         max_attractor_l=None, max_t=None):
    os.makedirs(output_dirpath, exist_ok=True)
    summaries, states = attractors_csv(db_conn, fixed_nodes, node_names, n_simulation_problems,
                                       max_attractor_l, max_t)
    _write(os.path.join(output_dirpath, aggregated_attractor_name + '_summaries.csv'), summaries)
    _write(os.path.join(output_dirpath, aggregated_attractor_name + 's.csv'), states)
