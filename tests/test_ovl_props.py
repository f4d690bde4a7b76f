"""CPU: the `.ovl` property checker (tests/ovl_props.py) that the GPU suite applies to outputs too large for a CPU comparison
(tests/test_gpu_wtzmo.py::test_cfg2_full_size_shard_properties).  It must accept what the reference / oracle print and reject
records that are wrong in any of the ways it claims to detect."""
import subprocess

import pytest

from ovl_props import check_ovl, load_fasta


@pytest.fixture(scope="module")
def sw_case(tmp_path_factory, gen_reads, oracle_bin):
    d = tmp_path_factory.mktemp("props")
    fa, ovl = str(d / "r.fa"), str(d / "r.ovl")
    subprocess.run([gen_reads, "-n", "300", "-L", "5000", "-G", "70000", "-s", "4", "-o", fa], check=True)
    subprocess.run([oracle_bin, "-t", "1", "-i", fa, "-f", "-o", ovl, "-k", "16", "-s", "200", "-m", "0.6"], check=True, stderr=subprocess.DEVNULL)
    return load_fasta(fa), ovl, d


def test_checker_accepts_reference_style_output(sw_case, gen_reads, oracle_bin, tmp_path):
    reads, ovl, _ = sw_case
    n, walked, cols = check_ovl(reads, ovl)
    assert n == walked > 100 and cols > 500000
    n2, walked2, _ = check_ovl(reads, ovl, walk_every=7)
    assert n2 == n and walked2 == (n + 6) // 7
    # dot-matrix records: coordinates only, "0M" in the CIGAR column
    fa, out = str(tmp_path / "o.fa"), str(tmp_path / "o.ovl")
    subprocess.run([gen_reads, "-n", "200", "-L", "4000", "-G", "60000", "-s", "7", "-m", "ont", "-o", fa], check=True)
    subprocess.run([oracle_bin, "-t", "1", "-i", fa, "-f", "-o", out, "-k", "16", "-z", "10", "-Z", "16", "-U", "-1", "-m", "0.1", "-A", "1000"], check=True, stderr=subprocess.DEVNULL)
    n, walked, cols = check_ovl(load_fasta(fa), out, dot_matrix=True)
    assert n > 30 and walked == 0 and cols > 0


def _edit(ovl, dst, line_no, fn):
    lines = open(ovl, "rb").read().split(b"\n")
    c = lines[line_no].split(b"\t")
    fn(c)
    lines[line_no] = b"\t".join(c)
    open(dst, "wb").write(b"\n".join(lines))


@pytest.mark.parametrize("what", ["shifted_coordinate", "wrong_count", "cigar_op_swapped", "duplicate_pair", "strand_flipped", "length", "identity_text"])
def test_checker_rejects_corrupted_records(sw_case, what):
    reads, ovl, d = sw_case
    bad = str(d / ("bad_%s.ovl" % what))
    if what == "shifted_coordinate":       # window moved by one base: spans still add up, the base-level walk does not
        def fn(c):
            c[3] = b"%d" % (int(c[3]) + 1); c[4] = b"%d" % (int(c[4]) + 1)
        ln = next(i for i, l in enumerate(open(ovl, "rb")) if int(l.split(b"\t")[4]) + 1 <= int(l.split(b"\t")[2]))
        _edit(ovl, bad, ln, fn)
    elif what == "wrong_count":
        def fn(c):
            c[12] = b"%d" % (int(c[12]) - 1); c[13] = b"%d" % (int(c[13]) + 1)
        _edit(ovl, bad, 3, fn)
    elif what == "cigar_op_swapped":
        def fn(c):
            c[16] = c[16].replace(b"I", b"#", 1).replace(b"D", b"I", 1).replace(b"#", b"D", 1)
        _edit(ovl, bad, 5, fn)
    elif what == "duplicate_pair":
        lines = open(ovl, "rb").read().split(b"\n")
        open(bad, "wb").write(b"\n".join(lines[:10] + [lines[2]] + lines[10:]))
    elif what == "strand_flipped":
        def fn(c):
            c[6] = b"-" if c[6] == b"+" else b"+"
        _edit(ovl, bad, 7, fn)
    elif what == "length":
        def fn(c):
            c[7] = b"%d" % (int(c[7]) + 1)
        _edit(ovl, bad, 1, fn)
    else:
        def fn(c):
            c[11] = b"0.999"
        _edit(ovl, bad, 2, fn)
    with pytest.raises(AssertionError):
        check_ovl(reads, bad)
