"""profiles/r2_traffic.json from committed ncu raw CSV exports: dram bytes read + written per launch of the fused kernel.

    ncu -i <capture>.ncu-rep --page raw --csv > profiles/<name>_raw.csv
    python tools/ncu_traffic.py f16=profiles/r2_affinity_idx_f16_raw.csv split3=profiles/r2_affinity_idx_split3_raw.csv

bench.py reads the file for roofline.traffic (the largest launch of each capture = the R = 9 launch)."""
import csv
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent


def main():
    out = {}
    for arg in sys.argv[1:]:
        key, path = arg.split('=', 1)
        rows = list(csv.reader(open(path)))
        hdr = rows[0]
        i_r, i_w, i_t = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('gpu__time_duration.sum')
        i_ur, i_uw = rows[1][i_r], rows[1][i_w]
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        best = max(rows[2:], key=lambda r: float(r[i_t]))
        rd, wr = float(best[i_r]) * scale[i_ur], float(best[i_w]) * scale[i_uw]
        out[key] = {'dram_bytes_per_launch': rd + wr, 'read': rd, 'written': wr, 'capture': str(Path(path).relative_to(REPO) if Path(path).is_absolute() else path),
                    'kernel': best[hdr.index('Kernel Name')][:80], 'duration_us_under_ncu': float(best[i_t]) / (1e3 if rows[1][i_t] == 'ns' else 1.0)}
    (REPO / 'profiles' / 'r2_traffic.json').write_text(json.dumps(out, indent=1) + '\n')
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
