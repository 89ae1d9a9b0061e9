"""Vectorised export writers against a straightforward per-atom formatter and the PDB / V2000 column rules."""
import os
import time

import numpy as np
import pytest

from diffbindfr_b200 import export, synth


def _naive_pdb(aatype, mask, a14):
    from diffbindfr_b200.constants import RESTYPES
    out, serial = [], 1
    for r, t in enumerate(aatype):
        res3 = export.RESNAME3[RESTYPES[int(t)]]
        for a, name in enumerate(export.ATOM14_NAMES[res3]):
            if not mask[r, a]:
                continue
            x, y, z = a14[r, a]
            an = (" " + name).ljust(4) if len(name) < 4 else name
            out.append("ATOM  %5d %s %s A%4d    %8.3f%8.3f%8.3f  1.00  0.00          %2s  \n" % (serial, an, res3, r + 1, x, y, z, name[0]))
            serial += 1
    return out


def test_pdb_and_sdf_templates_roundtrip_and_match_naive_formatter(tmp_path):
    rng = np.random.default_rng(0)
    s = synth.make_sample(rng, 24, 19)
    P = 5
    a14 = np.stack([s["atom14_position"] + rng.normal(scale=0.3, size=s["atom14_position"].shape) * s["atom14_mask"][..., None] for _ in range(P)])
    lig = np.stack([s["lig_pos"] + rng.normal(scale=1.0, size=s["lig_pos"].shape) for _ in range(P)])
    pt = export.PdbTemplate(s["sequence"], s["atom14_mask"])
    texts = pt.render(a14)
    assert len(texts) == P
    for p in range(P):
        lines = texts[p].splitlines(keepends=True)
        atom_lines = [ln for ln in lines if ln.startswith("ATOM")]
        assert atom_lines == _naive_pdb(s["sequence"], s["atom14_mask"], a14[p])
        assert all(len(ln) == 81 for ln in atom_lines) and lines[-1] == "END\n" and lines[-2].startswith("TER")
        xyz = export.parse_pdb_coords(texts[p])
        assert np.abs(xyz - a14[p][s["atom14_mask"].astype(bool)]).max() <= 5.1e-4
    ei = s["lig_edge_index"]
    bonds = ei[:, ei[0] < ei[1]].T
    st = export.SdfTemplate(["C"] * lig.shape[1], bonds, name="lig0")
    recs = st.render(lig)
    for p in range(P):
        ls = recs[p].splitlines()
        assert ls[0] == "lig0" and ls[3].endswith("V2000") and int(ls[3][:3]) == lig.shape[1] and int(ls[3][3:6]) == len(bonds)
        assert ls[-1] == "$$$$" and ls[-2] == "M  END"
        assert np.abs(export.parse_sdf_coords(recs[p]) - lig[p]).max() <= 5.1e-5
        assert [(int(l[0:3]) - 1, int(l[3:6]) - 1) for l in ls[4 + lig.shape[1]:4 + lig.shape[1] + len(bonds)]] == [tuple(b) for b in bonds.tolist()]
    st2 = export.SdfTemplate.from_molblock(recs[0])             # topology of an existing mol block is kept verbatim
    assert st2.render(lig[1])[0] == recs[1]
    paths = export.export_poses(str(tmp_path), "cplx", st, pt, lig, a14, pocket_center=np.array([10.0, -5.0, 3.0]))
    assert [os.path.basename(os.path.dirname(p["docked_lig"])) for p in paths] == [f"sample_{k + 1}" for k in range(P)]
    back = export.parse_sdf_coords(open(paths[2]["docked_lig"]).read())
    assert np.abs(back - (lig[2] + np.array([10.0, -5.0, 3.0]))).max() <= 5.1e-5
    pb = export.parse_pdb_coords(open(paths[2]["protein_pdb"]).read())
    assert np.abs(pb - (a14[2][s["atom14_mask"].astype(bool)] + np.array([10.0, -5.0, 3.0]))).max() <= 5.1e-4


def test_export_throughput_40_poses():
    """The reference needs ~32 s for the 40 poses of one complex; the templates format them in well under a second."""
    rng = np.random.default_rng(1)
    s = synth.make_sample(rng, 105, 35, 12.0 * (105 / 36.0) ** (1.0 / 3.0))
    a14 = np.repeat(s["atom14_position"][None], 40, 0); lig = np.repeat(s["lig_pos"][None], 40, 0)
    pt = export.PdbTemplate(s["sequence"], s["atom14_mask"]); st = export.SdfTemplate(["C"] * 35, np.zeros((0, 2), int))
    t0 = time.perf_counter()
    a, b = pt.render(a14), st.render(lig)
    dt = time.perf_counter() - t0
    assert len(a) == 40 and len(b) == 40 and dt < 2.0, dt


def test_error_corrected_export_carries_the_smina_tags(tmp_path):
    """``lig_final_ec.sdf`` with ``minimizedAffinity`` / ``minimizedRMSD`` SD items, the file ``error_corrector`` leaves per pose."""
    from diffbindfr_b200 import export
    t = export.SdfTemplate(["C", "N", "O"], np.array([[0, 1], [1, 2]]), [1, 2])
    poses = np.arange(18, dtype=np.float64).reshape(2, 3, 3) / 7.0
    paths = export.export_corrected(str(tmp_path), "cx", t, poses, [-9.43802, 1.5], pocket_center=np.array([1.0, 2.0, 3.0]), rmsd_to_start=[0.4, 0.1])
    assert [p.split("/")[-2:] for p in paths] == [["sample_1", "lig_final_ec.sdf"], ["sample_2", "lig_final_ec.sdf"]]
    assert export.read_sd_tag(paths[0], "minimizedAffinity") == pytest.approx(-9.43802) and export.read_sd_tag(paths[1], "minimizedRMSD") == pytest.approx(0.1)
    txt = open(paths[1]).read()
    assert txt.rstrip().endswith("$$$$") and txt.count("M  END") == 1
    assert np.allclose(export.parse_sdf_coords(txt), poses[1] + np.array([1.0, 2.0, 3.0]), atol=1e-4)
