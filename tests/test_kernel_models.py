"""CPU: sequential models of the exact shortcuts the kernels take (DESIGN.md §2 "Exact shortcuts"), each checked against the
literal algorithm of the reference on random inputs.  The GPU parity tests check the kernels against the oracle; these check
the arithmetic of the shortcuts themselves, without a GPU."""
import os
import subprocess

import parity_util as pu

BIN = os.path.join(pu.ROOT, "tests", "cpu_models", "_bin")


def test_run_compressed_dp_and_run_wise_walk_equal_the_literal_algorithm(built):
    """chainRunDpKernel / chainFillKernel / chainWalkKernel<true> (overlap.cu) vs overlap.cpp:277-323 and :338-383: scores, back
    pointers and the (start, first match, length) of every chain, on 60 k random pairs (diagonal runs, indels, jumps beyond
    maxJump, noise, ties; either sort axis); the binary search for a run's first valid match and the presorted shortcut are hit."""
    r = subprocess.run([os.path.join(BIN, "rundp_check"), "60000", "19"], stdout=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout
    stats = dict(kv.split("=") for kv in r.stdout.split()[3:])
    assert int(stats["walkBacks"]) > 1000 and int(stats["presorted"]) > 50 and int(stats["heads"]) > 100000


def test_bounded_band_pruned_wavefronts_are_exact_below_the_limit(built):
    """wfaSteps / wfaEditDistance / dropLimit (editdist.cu) vs the O(nm) edit distance: exact for distances below the limit,
    '>= limit' otherwise, for limits around the true distance; dropLimit is the smallest distance failing the float test."""
    r = subprocess.run([os.path.join(BIN, "wfa_check"), "12000", "23"], stdout=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout


def test_segmented_radix_sort_model_is_the_stable_sort_by_target(built):
    """segRadixSortKernel (overlap.cu): per-warp chunks, (warp, digit) offsets, match.any ranking, next-pass counts gathered
    while scattering == std::stable_sort by extId for 1..3 passes and every segment shape; and for segments without an
    (extId, curPos) tie that order is what std::sort by (extId, curPos) returns from any input order."""
    r = subprocess.run([os.path.join(BIN, "segsort_check"), "1200", "29"], stdout=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout
    assert int(r.stdout.split("tieFree=")[1]) > 100


def test_dense_kmer_class_index_is_a_bijection_onto_the_canonical_kmers(built):
    """denseIndexFromWindow / denseIndexOfPair / canonFromDenseIndex (kmer_math.cuh), the addressing of the counting kernels
    (count_index.cu): both strands of a k-mer share one index below the array size, the index maps back to Kmer::standardForm's
    canonical k-mer, classes never collide (exhaustive for k <= 9, sampled up to k = 17), and the multi-GPU owner / slot split
    (index % N, index / N) is a bijection with 32-bit slots."""
    r = subprocess.run([os.path.join(BIN, "dense_check")], stdout=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout


def test_device_logf_is_glibc_logf_bit_for_bit(built):
    """glibcLogf / kmerDivergence (glibc_logf.cuh, used by divergenceKernel in overlap.cu) vs the container's logf — the function
    behind std::log(float) at overlap.cpp:423: identical bits for every 97th positive normal float, every float of [1, 4), and
    for the whole seqDivergence expression on 4 M random records (`logf_check 1` compares all 2.13e9 positive normal floats)."""
    r = subprocess.run([os.path.join(BIN, "logf_check"), "97"], stdout=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout
    stats = dict(kv.split("=") for kv in r.stdout.split()[1:])
    assert int(stats["n"]) > 30000000 and int(stats["degenerate"]) > 1000
