"""Print the handful of numbers of a bench.py JSON line that one looks at first."""
import json
import sys

j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
d = j.get("device") or {}
print(f"verified={j.get('verified')} stall={j.get('stall_ms_per_handoff', -1):.0f} ms e2e={(j.get('e2e') or {}).get('value') or -1:.1f} GB/s "
      f"iter/s={j.get('iter_per_s', -1):.2f} value={j.get('value')}")
if d:
    w = d.get("wall_ms_mean") or {}
    print(f"  evict {d.get('evict_GBps')} fetch {d.get('fetch_GBps')} scan {d.get('scan_GBps')}; per step: out {d['bytes_evicted'] / 1e9 / j['steps']:.1f} GB "
          f"in {d['bytes_fetched'] / 1e9 / j['steps']:.1f} GB clean {d['bytes_skipped_clean'] / 1e9 / j['steps']:.1f} GB; wall evict {w.get('evict')} fetch {w.get('fetch')} "
          f"map {d.get('map_ms_mean')} wait {d.get('wait_ms_mean')}; excess {d.get('link_bytes_over_algorithmic')}")
print("  gaps", [round(g) for g in j.get("first_iter_gap_ms", [])], j.get("analysis_error") or "", j.get("error") or "")
for k, v in (j.get("configs") or {}).items():
    print("  config", k, {kk: v.get(kk) for kk in ("verified", "stall_ms_per_handoff", "e2e_GBps", "iter_per_s", "error")})
if "same_scale" in j:
    s = j["same_scale"]
    print("  same_scale", {kk: s.get(kk) for kk in ("verified", "hbm_fraction_used", "stall_ms_per_handoff", "e2e_GBps", "iter_per_s", "error")})
