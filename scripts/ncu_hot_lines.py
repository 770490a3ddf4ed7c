"""Top source lines of an ncu capture (needs --import-source on and -lineinfo):
    python scripts/ncu_hot_lines.py <file.ncu-rep> [n]
Prints, per source line, executed warp instructions, share of the kernel, stall samples and average active threads."""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows, fname, hdr = [], None, None
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].strip().isdigit():
        # source text may hold unescaped quotes (asm strings): address the numeric columns from the right
        def num(k):
            v = r[hdr.index(k) - len(hdr)]
            try:
                return float(v)
            except ValueError:
                return 0.0
        rows.append((fname, int(r[0]), r[1], num("Instructions Executed"), num("# Samples"), num("Thread Instructions Executed")))
tot = sum(x[3] for x in rows) or 1.0
tots = sum(x[4] for x in rows) or 1.0
print("total warp instructions %.4g, stall samples %.4g" % (tot, tots))
for f, ln, src, ins, smp, tins in sorted(rows, key=lambda x: -x[3])[:top]:
    print("%5.1f%% inst %5.1f%% smp  lanes %4.1f  %s:%d  %s" % (100 * ins / tot, 100 * smp / tots, tins / ins if ins else 0, f, ln, src.strip()[:110]))
