import os, sys, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
import parity_util as pu
import test_gpu_host_mirror as t
tmp = tempfile.mkdtemp()
a = pu.simulate(os.path.join(tmp, "a.fasta"), genome_len=120000, coverage=6, mean_len=12000, shape=20, error=0.002, seed=51)
b = pu.simulate(os.path.join(tmp, "b.fasta"), genome_len=120000, coverage=8, mean_len=6000, shape=4, error=0.03, seed=51)
opts = ["--query-reads", b, "--all-ext", "--both-strands", "--no-estimate", "--min-overlap", "500"]
ref = pu.run_oracle(a, t.HIFI, os.path.join(tmp, "ref"), extra=opts)
got = t.run_mirror(a, t.HIFI, os.path.join(tmp, "gpu"), extra=opts)
print(ref["overlaps"], got["overlaps"])
ra = open(os.path.join(tmp, "ref.ovlp")).read().splitlines(); ga = set(open(os.path.join(tmp, "gpu.ovlp")).read().splitlines())
miss = [l for l in ra if l not in ga]
print(len(miss), "reference lines missing on the device, e.g.")
for l in miss[:12]:
    print(l)
print("header/sample of ref:", ra[:3])
