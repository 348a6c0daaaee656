#!/usr/bin/env python3
"""tests/golden/week171_subset.rnx: the header and the Galileo I/NAV records (data source 517, the only
ones the reference keeps, src/rinex.cpp:218) with clock epochs in hours 00-01 and 11-13 of 2021-06-20 from
the reference's rinex_files/week171.rnx -- public IGS broadcast-ephemeris DATA (merged by gfzrnx, see
its header), not reference source code.  The two golden scenarios (tests/golden/*_recs.npz) only ever
match records from those hours, and the reference picks the FIRST record within an hour of the
requested time in file order, so the subset yields the same records as the full file (checked by
tests/test_host_scenario.py against the trace-derived goldens).  Build container only: reads
/root/reference."""
import sys
from pathlib import Path

SRC = Path("/root/reference/rinex_files/week171.rnx")
DST = Path(__file__).resolve().parent.parent / "tests" / "golden" / "week171_subset.rnx"

lines = SRC.read_text().splitlines(keepends=True)
out, i = [], 0
while i < len(lines):
    out.append(lines[i])
    i += 1
    if "END OF HEADER" in out[-1]:
        break
kept = 0
while i < len(lines):
    ln = lines[i]
    if ln[0] != " ":                       # first line of a record: 8 lines for Galileo, GPS 8, GLONASS 4 ...
        n = 4 if ln[0] in "RS" else 8
        if (ln[0] == "E" and ln[4:14] == "2021 06 20" and int(ln[15:17]) in (0, 1, 11, 12, 13)
                and float(lines[i + 5][23:42].replace("D", "E")) == 517.0):
            out.extend(lines[i:i + n])
            kept += 1
        i += n
    else:
        i += 1
DST.write_text("".join(out))
print(DST, kept, "records", DST.stat().st_size, "bytes", file=sys.stderr)
