"""Where do a GEMM kernel's warp-stall samples land?  python tools/ncu_src_roles.py <file.ncu-rep> <kernel substring> [nth]

Reads the SASS source page of a `ncu --set full --import-source on` capture and splits the samples of one gemm_kernel
launch by warp role (TMA producer, MMA issuer, epilogue), using the role's marker instructions (UTMALDG, UTCHMMA, LDTM)
as range boundaries, then lists the hottest instructions.  Needs `ncu` on PATH; no GPU."""
import csv
import subprocess
import sys

rep, name = sys.argv[1], sys.argv[2]
nth = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks = [b for b in out.split('"Kernel Name",') if name in b.split("\n", 1)[0]]
blk = blocks[nth]
rows = list(csv.reader(blk.splitlines()[1:]))
h = rows[0]
ai, ei = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
rows = [r for r in rows[1:] if len(r) > ai]
tot = sum(int(r[ai]) for r in rows)
first = {k: next((i for i, r in enumerate(rows) if k in r[1] and int(r[ei]) > 0), None) for k in ("UTMALDG", "LDTM", "UTCHMMA")}
print(f"kernel: {blk.splitlines()[0][:100]}  samples {tot}, {len(rows)} SASS instructions, first markers {first}")
waits = [(i, int(rows[i + 1][ai]) + int(r[ai])) for i, r in enumerate(rows[:-1]) if "TRYWAIT" in r[1]]
for i, s in waits:
    if s > 0.01 * tot:
        print(f"  mbarrier wait at #{i}: {100 * s / tot:5.1f} % of samples   {rows[i][1].strip()[:70]}")
print("hottest instructions:")
for i, r in sorted(enumerate(rows), key=lambda t: -int(t[1][ai]))[:14]:
    print(f"  #{i:5d} {100 * int(r[ai]) / tot:5.1f} %  exec {r[ei]:>9}  {r[1].strip()[:80]}")
