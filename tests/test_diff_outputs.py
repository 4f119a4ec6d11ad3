"""integration/diff_outputs.py (the token diff of the side-by-side runs) on synthetic text: tolerances, skipped lines, reference-noisy
lines, structure-only scripts — CPU only."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("diff_outputs", os.path.join(ROOT, "integration", "diff_outputs.py"))
D = importlib.util.module_from_spec(spec); spec.loader.exec_module(D)


def test_numbers_within_print_precision_match_and_beyond_do_not():
    ref = "loss=2.37784 \nvector[3] = { +0.1445 +0.1889 -0.2433 }\n"
    ok_ = "loss=2.37790 \nvector[3] = { +0.1446 +0.1889 -0.2433 }\n"
    bad = "loss=2.38784 \nvector[3] = { +0.1445 +0.1889 -0.2433 }\n"
    assert D.compare(ref, ok_, None, D.RTOL)[0] == []
    assert len(D.compare(ref, bad, None, D.RTOL)[0]) == 1
    assert D.compare(ref, bad, None, 2e-2)[0] == []                      # a LOOSE script's bar


def test_words_must_match_and_glued_numbers_split():
    assert D.compare("relu [ 2, 8, 8, 2] 0.05-0.21 0.07", "relu [ 2, 8, 8, 2] 0.05-0.21 0.07", None, D.RTOL)[0] == []
    assert D.compare("relu [ 2, 8, 8, 2] 0.05-0.21", "tanh [ 2, 8, 8, 2] 0.05-0.21", None, D.RTOL)[0] != []
    assert D.toks(" 0.05-0.21 0.07") == ["0.05", "-0.21", "0.07"]


def test_pointer_banner_and_timestamp_lines_are_ignored():
    ref = "\\ TLSF: ostore=0x7fa0ee000000, alloc=0x80000000\n  0.00:  2> linear  [ 1, 1, 3, 1] p= 1.000\n"
    new = "\\ TLSF: ostore=0x7fec56000000, alloc=0x80000000\n  1.00:  2> linear  [ 1, 1, 3, 1] p= 1.000\n"
    assert D.compare(ref, new, None, D.RTOL)[0] == []


def test_lines_the_reference_does_not_reproduce_are_structure_only():
    ref, ref2 = "w= { +0.0773 -0.4053 }\nverify { 6 13 20 }\n", "w= { -0.3104 -0.2184 }\nverify { 6 13 20 }\n"
    new_ok, new_bad = "w= { +0.9999 +0.0306 }\nverify { 6 13 20 }\n", "w= { +0.9999 +0.0306 }\nverify { 6 13 21 }\n"
    bad, nnum, noisy, *_ = D.compare(ref, new_ok, ref2, D.RTOL)
    assert bad == [] and noisy == 2 and nnum == 5
    assert D.compare(ref, new_bad, ref2, D.RTOL)[0] != []               # the stable line is still held to the bar
    assert D.compare(ref, new_bad, None, D.RTOL, all_free=True)[0] == []   # STRUCTURE_ONLY scripts
