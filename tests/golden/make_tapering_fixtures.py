"""Runs the REFERENCE'S OWN Z2Symmetries (qiskit/aqua/operators/legacy/weighted_pauli_operator.py:
960-1330; kernel_F2 in legacy/common.py) -- two_qubit_reduction on parity-mapped molecular
Hamiltonians, find_Z2_symmetries and taper (all sectors) -- and records the results in
tests/golden/reference_tapering.json.

    python tests/golden/make_tapering_fixtures.py        (needs /root/reference; CPU only)

Same arrangement as make_reference_fixtures.py (terra stubbed, package __init__ files bypassed).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_fixtures as mrf  # noqa: E402
import terra_stub  # noqa: E402


def cplx(v):
    v = complex(v)
    return [v.real, v.imag]


def plist(op):
    return [[cplx(w), p.to_label()] for w, p in op.paulis]


def main():
    R = mrf.load_reference()
    WPO, Pauli = R['WPO'], terra_stub.Pauli
    Z2 = sys.modules['qiskit.aqua.operators.legacy.weighted_pauli_operator'].Z2Symmetries
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import pauli_oracle as po
    chem = json.load(open(os.path.join(HERE, 'chemistry_integrals.json')))
    out = {'note': 'outputs of the reference Z2Symmetries (terra Pauli stubbed); see make_tapering_fixtures.py',
           'two_qubit_reduction': [], 'find_and_taper': []}
    # ---- parity mapping + two-qubit reduction: H2 (the reference's G1 operator) and LiH ---------
    for name, key, particles in (('h2', 'G10_h2_fcidump', 2), ('lih', 'G11_lih_npz', [2, 2])):
        h1 = po.onee_to_spin(np.array(chem[key]['mo_onee']))
        h2 = po.twoe_to_spin(np.array(chem[key]['mo_eri']))
        op = R['FermionicOperator'](h1, h2).mapping('parity')
        red = Z2.two_qubit_reduction(op, particles)
        rec = {'name': name, 'num_particles': particles, 'parity_terms': len(op.paulis), 'num_qubits': int(red.num_qubits),
               'reduced': plist(red) if name == 'h2' else None, 'reduced_terms': len(red.paulis),
               'tapering_values': [int(v) for v in red.z2_symmetries.tapering_values],
               'sq_list': [int(v) for v in red.z2_symmetries.sq_list]}
        if name == 'lih':
            d = {p.to_label(): complex(w) for w, p in red.paulis}
            labs = sorted(d)
            rec['reduced_sample'] = [[cplx(d[l]), l] for l in labs[::7]]
            rec['reduced_weight_sum'] = cplx(sum(d.values()))
            rec['reduced_abs_sum'] = float(sum(abs(v) for v in d.values()))
        out['two_qubit_reduction'].append(rec)
    # ---- find_Z2_symmetries + taper in every sector -----------------------------------------------
    cases = []
    h1 = po.onee_to_spin(np.array(chem['G10_h2_fcidump']['mo_onee']))
    h2 = po.twoe_to_spin(np.array(chem['G10_h2_fcidump']['mo_eri']))
    for mp in ('jordan_wigner', 'parity', 'bravyi_kitaev'):
        op = R['FermionicOperator'](h1, h2).mapping(mp)
        cases.append(('h2_' + mp, [(complex(w), p.to_label()) for w, p in op.paulis]))
    cases.append(('heisenberg4', [(1.0, 'XXII'), (1.0, 'YYII'), (1.0, 'ZZII'), (0.5, 'IXXI'), (0.5, 'IYYI'), (0.5, 'IZZI'),
                                  (0.25, 'IIXX'), (0.25, 'IIYY'), (0.25, 'IIZZ')]))
    cases.append(('tfim5', [(-1.0, 'ZZIII'), (-1.0, 'IZZII'), (-1.0, 'IIZZI'), (-1.0, 'IIIZZ'),
                            (0.7, 'XIIII'), (0.7, 'IXIII'), (0.7, 'IIXII'), (0.7, 'IIIXI'), (0.7, 'IIIIX')]))
    cases.append(('mixed3', [(0.3, 'XYZ'), (1.1, 'YXZ'), (-0.4, 'ZZI'), (0.9, 'IIZ')]))
    cases.append(('nosym2', [(1.0, 'XI'), (1.0, 'ZI'), (1.0, 'IX'), (1.0, 'IZ'), (0.5, 'YY')]))
    for name, terms in cases:
        op = WPO(paulis=[[w, Pauli(l)] for w, l in terms])
        z2 = Z2.find_Z2_symmetries(op)
        rec = {'name': name, 'terms': [[cplx(w), l] for w, l in terms],
               'symmetries': [s.to_label() for s in z2.symmetries], 'sq_paulis': [s.to_label() for s in z2.sq_paulis],
               'sq_list': [int(v) for v in z2.sq_list], 'sectors': None}
        if not z2.is_empty():
            tapered = z2.taper(op)
            rec['sectors'] = [{'tapering_values': [int(v) for v in t.z2_symmetries.tapering_values], 'paulis': plist(t),
                               'num_qubits': int(t.num_qubits) if not t.is_empty() else None} for t in tapered]
        out['find_and_taper'].append(rec)
    path = os.path.join(HERE, 'reference_tapering.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=0)
    print('wrote', path, os.path.getsize(path), 'bytes')
    print(out['two_qubit_reduction'][0]['reduced'])
    for r in out['find_and_taper']:
        print(r['name'], r['symmetries'], r['sq_paulis'], r['sq_list'])


if __name__ == '__main__':
    main()
