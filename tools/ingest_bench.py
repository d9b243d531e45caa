"""Dev / measurement tool for SURVEY 8(f) N2: rows per second of the three CSV -> structure-of-arrays routes on a synthetic
file in the modern-ship format (quoted ids, NA fields): the reference's way (one pandas parse PER SHIP, timed on a few
ships and scaled), ingest.read_csv_fleet (one pandas parse) and ingest.read_csv_fleet_device (bytes parsed on the GPU).
usage: python tools/ingest_bench.py [--rows 20000000] [--ships 20000] [--file path]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=20_000_000)
ap.add_argument("--ships", type=int, default=20_000)
ap.add_argument("--file", default="/tmp/ste_ingest_bench.csv")
ap.add_argument("--skip-host", action="store_true")
a = ap.parse_args()

if not os.path.exists(a.file) or os.path.getsize(a.file) < 40 * a.rows:
    t0 = time.perf_counter()
    rng = np.random.default_rng(1)
    with open(a.file, "w") as fh:
        fh.write('"","yr","mo","dy","hr","dck","id","lat","lon","w","d"\n')
        chunk = 1_000_000
        hour = 0
        for lo in range(0, a.rows, chunk):
            n = min(chunk, a.rows - lo)
            k = rng.integers(0, a.ships, n)
            hrs = hour + np.cumsum(rng.integers(0, 2, n))          # non-decreasing time through the file
            hour = int(hrs[-1])
            day = hrs // 24
            yr, doy = 2021 + day // 336, day % 336                 # 12 months of 28 days: every date is valid
            mo, dy, hr = doy // 28 + 1, doy % 28 + 1, hrs % 24
            lat = np.round(rng.uniform(-60, 60, n), 3)
            lon = np.round(rng.uniform(0, 359, n), 3)
            idx = np.arange(lo, lo + n)
            fh.write("\n".join(f'"{i}",{y},{m},{d},{h},992,"S{s:06d}",{la},{lo_},NA,NA'
                               for i, y, m, d, h, s, la, lo_ in zip(idx, yr, mo, dy, hr, k, lat, lon)) + "\n")
    print(json.dumps({"generated": a.file, "rows": a.rows, "GB": os.path.getsize(a.file) / 1e9, "s": time.perf_counter() - t0}), flush=True)

from ship_track_estimators_b200.ingest import read_csv_fleet, read_csv_fleet_device
kw = dict(id_col="id", lat_col="lat", lon_col="lon")
t0 = time.perf_counter(); dev = read_csv_fleet_device(a.file, **kw); t_dev = time.perf_counter() - t0   # includes CUDA context + first-touch
t0 = time.perf_counter(); dev = read_csv_fleet_device(a.file, **kw); t_dev2 = time.perf_counter() - t0
print(json.dumps({"route": "device (ste_csv_parse_rows)", "rows": dev.stats["rows"], "ships": dev.n_tracks, "first_call_s": t_dev, "s": t_dev2,
                  "rows_per_s": dev.stats["rows"] / t_dev2, "stages": dev.stats}), flush=True)
if not a.skip_host:
    t0 = time.perf_counter(); host = read_csv_fleet(a.file, **kw); t_host = time.perf_counter() - t0
    print(json.dumps({"route": "host, one pandas parse (read_csv_fleet)", "rows": dev.stats["rows"], "s": t_host, "rows_per_s": dev.stats["rows"] / t_host}), flush=True)
    h = dev.to_host()
    print(json.dumps({"identical": bool(h.ids == host.ids and np.array_equal(h.lon, host.lon) and np.array_equal(h.lat, host.lat)
                                        and np.array_equal(h.dts, host.dts) and np.array_equal(h.n_obs, host.n_obs))}), flush=True)
    from ship_track_estimators_b200.ship_track import ShipTrack
    t0 = time.perf_counter()
    for sid in host.ids[:2]:
        ShipTrack().read_csv(a.file, ship_id=sid, **kw)
    per_ship = (time.perf_counter() - t0) / 2
    print(json.dumps({"route": "reference's way: one parse per ship (ShipTrack.read_csv)", "s_per_ship": per_ship,
                      "extrapolated_s_all_ships": per_ship * host.n_tracks, "rows_per_s": dev.stats["rows"] / (per_ship * host.n_tracks)}), flush=True)
