"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump: samples per source line and the hottest SASS, per kernel."""
import csv, sys, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Function Name":
        fn = rows[i][1]; hdr = rows[i + 1]; i += 2
        body = []
        while i < len(rows) and not (rows[i] and rows[i][0] == "File Path"):
            if len(rows[i]) == len(hdr): body.append(rows[i])
            i += 1
        c = {h: k for k, h in enumerate(hdr)}
        src_col = [k for k, h in enumerate(hdr) if h == "Source"]
        samp = c.get("# Samples")
        if samp is None: continue
        # rows with a line number are CUDA-C lines; rows with an address are SASS
        tot = 0; by_line = []; sass = []
        for r in body:
            try: n = int(r[samp])
            except ValueError: continue
            if r[c["Address"]]: sass.append((n, r[src_col[1]] if len(src_col) > 1 else r[src_col[0]], r)); tot += n
            elif r[c["Line No"]]: by_line.append((n, r[c["Line No"]], r[src_col[0]].strip()[:150]))
        print("=" * 20, fn[:150]); print("total samples", tot)
        for n, ln, s in sorted(by_line, reverse=True)[:topn]:
            print("%7d %5.1f%%  L%-4s %s" % (n, 100.0 * n / max(tot, 1), ln, s))
        print("-- hottest SASS")
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        for n, s, r in sorted(sass, key=lambda x: -x[0])[:topn]:
            st = sorted(((int(r[c[h]] or 0), h) for h in stall_cols), reverse=True)[:2]
            print("%7d %5.1f%%  %-70s %s" % (n, 100.0 * n / max(tot, 1), s[:70], " ".join("%s=%d" % (h[6:], v) for v, h in st if v)))
    else:
        i += 1
