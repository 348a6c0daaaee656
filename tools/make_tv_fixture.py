#!/usr/bin/env python3
"""tests/golden/tv_pages.csv: a sample of the reference's tv/<date>/<svid>.csv files -- live-sky Galileo I/NAV pages
(`TOW,WN,SVID,240-bit hex`: even half page incl. tail, odd half page incl. tail), recorded DATA the reference ships for
its retired replay mode, not source code.  One 60 s cycle of word types (30 pages) of four satellites on three days.
tests/test_tv_pages.py holds the host page builder's CRC-24Q, channel coding and time-field layout, and the receiver
stand-in's decoder, against them.  Build container only: reads /root/reference."""
import sys
from pathlib import Path

SRC = Path("/root/reference/tv")
DST = Path(__file__).resolve().parent.parent / "tests" / "golden" / "tv_pages.csv"
rows = []
for day in ("11_DEC_2020_GST_08_00_01", "12_JAN_2021_GST_10_00_01", "20_FEB_2022_GST_08_00_01"):
    files = sorted((SRC / day).glob("*.csv"), key=lambda p: int(p.stem))
    files = [f for f in files if not f.read_text().splitlines()[0].split(",")[3].startswith("3F")][:4]   # skip satellites sending dummy pages only
    for f in files:
        for ln in f.read_text().splitlines()[:30]:
            rows.append(f"{day[:11]},{ln.strip()}")
DST.write_text("\n".join(rows) + "\n")
print(DST, len(rows), "pages", DST.stat().st_size, "bytes", file=sys.stderr)
