"""Pins oracle/setup_oracle.py (lte_pops, compute_collisions, v_broad restated on plain arrays) bit for bit to what
the unmodified reference produced: every column fixture holds the reference's nStar, C and vBroad; the inputs those
came from are in tests/golden/setup_inputs.npz (tests/golden/make_golden.py setup)."""
import numpy as np
import pytest

from helpers import load_golden, load_setup_inputs

FIXTURES = ['c1_falc_ca', 'c2_falc_cah', 'c1v_jitter_ca3', 'c2v_jitter_cah_0', 'c2v_jitter_cah_1', 'rf_k40p', 'rf_k10m',
            'stress_r10_d512']


@pytest.mark.parametrize('name', FIXTURES)
def test_lte_pops_collisions_vbroad_bitwise(name):
    from oracle import setup_oracle as so
    p, _ = load_golden(name)
    atoms, col = load_setup_inputs(name)
    assert np.array_equal(col['temperature'], p['temperature'])
    lvl = np.concatenate([[0], np.cumsum(p['Nlevel'])]).astype(int)
    g2 = np.concatenate([[0], np.cumsum(np.asarray(p['Nlevel'], dtype=int) ** 2)]).astype(int)
    for a, nm in enumerate(str(s).strip().upper() for s in p['atom_names']):
        A = atoms[nm]
        nTotal = A['abundance'] * col['nHTot']
        assert np.array_equal(nTotal, p['nTotal'][a])
        nStar = so.lte_pops(A['E_SI'], A['g'], A['stage'], col['temperature'], col['ne'], nTotal)
        assert np.array_equal(nStar, p['nStar'][lvl[a]:lvl[a + 1]]), (name, nm)
        C = so.compute_collisions(A['E_SI'], A['g'], A['coll'], A['coll_T'], A['coll_rates'], col['temperature'],
                                  col['ne'], nStar)
        NL = int(p['Nlevel'][a])
        assert np.array_equal(C.reshape(NL * NL, -1), p['C'][g2[a]:g2[a + 1]]), (name, nm)
        assert np.array_equal(so.v_broad(A['weight'], col['temperature'], p['vturb']), p['vBroad'][a]), (name, nm)
