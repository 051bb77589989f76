"""CPU stress on the fake driver: N multi-threaded clients (tests/apps/mt_app.c in io mode: launches,
device copies, and host copies that are served by the device path or from the backing copy
depending on where the lock is) under one scheduler with TQ = 1 s, with shared pools that are
tight or outright too small and both eviction policies.  Every client must finish with
mismatches=0.  Usage: python tools/stress_fake.py"""
import sys, subprocess, tempfile, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent / "tests"))
from nvs_testlib import ORACLE, Daemon, fake_env, preload
def run(n_clients, pool_mib, policy, seconds, total=200, threads=3, mib=12):
    tmp=pathlib.Path(tempfile.mkdtemp()); sd=tmp/'nvs'; sd.mkdir()
    d=Daemon("ours", sd, log_path=tmp/'sched.log')
    try:
        d.ctl("-T","1")
        procs=[]
        for i in range(n_clients):
            env=fake_env(total_mib=total, ledger=tmp/'ledger', extra={"NVSHARE_HOST_ARENA_MIB":64,"NVSHARE_CHUNK_MIB":8,"NVSHARE_BATCH_MIB":32,
                 "NVSHARE_DEBUG":1,"NVSHARE_SOCK_DIR":sd,"NVSHARE_POOL_MIB":pool_mib,"NVSHARE_EVICT_POLICY":policy,"NVSHARE_OOM_WAIT_MS":15000})
            env["LD_PRELOAD"]=preload("ours")
            procs.append(subprocess.Popen([str(ORACLE/"mt_app"),str(mib),str(seconds),str(i+1),str(threads),"1"],env=env,stdout=subprocess.PIPE,stderr=subprocess.PIPE,text=True))
        outs=[p.communicate(timeout=300) for p in procs]
        ok=all(p.returncode==0 and "RESULT PASS" in o for p,(o,e) in zip(procs,outs))
        drops=d.read_log().count("Sent DROP_LOCK")
        served=sum(e.count("served from the backing copy") for o,e in outs)
        print(f"clients={n_clients} pool={pool_mib} policy={policy}: ok={ok} drops={drops} served={served}", [o.strip().splitlines()[-1] if o.strip() else e[-200:] for o,e in outs])
        return ok
    finally:
        d.stop()
# per client: threads x (2 x mib + ~3) MiB
allok=True
for args in [(3,128,"need",8),(3,128,"all",8),(2,64,"all",8),(4,192,"need",8),(3,64,"need",8)]:
    allok &= run(*args)
print("ALL OK" if allok else "FAILURES")
