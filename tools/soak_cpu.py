#!/usr/bin/env python3
"""CPU soak (test infrastructure): fuzzer-made streams of every layout, some truncated or with a flipped bit, through
the kernel bodies (tests/hostsim), the plain-C oracle and -- for streams that decode without an error status -- the
compiled reference; prints every mismatch.  With "rom" as the second argument: seeded ROM-playback scenarios (all four OS
versions, error streams, software 1.05) through the sequencer + kernel bodies against the reference decoder.
usage: tools/soak_cpu.py [seconds=1500] [streams|rom|encode]"""
import os
import sys
import time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dcsfuzz, simutil
from oracle import orc, ref
if len(sys.argv) > 2 and sys.argv[2] == "encode":
    # the encoder's kernel bodies (hostsim) against the reference encoder fed the same framing: seeded clips and parameters
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import simutil
    import encode_soak_cases as esc
    budget = float(sys.argv[1])
    t0, n, bad, k = time.time(), 0, 0, 0
    while time.time() - t0 < budget:
        cp = [esc.make((97, k * 40 + j)) for j in range(40)]
        cp = [(x[:6000], p if p[0] >= 0 and p[1] >= 0 else (0, 0) + p[2:]) for x, p in cp]       # (the wildcard is host logic of the product)
        got = simutil.encode_streams([c for c, _ in cp], [p for _, p in cp])
        for (x, p), g in zip(cp, got):
            w = ref.encode_framed(x, p[0], p[1], p[2], p[3], p[4], p[5], fmt=p[6])[0]
            n += 1
            if g != w:
                bad += 1
                print("MISMATCH clip", k, p)
        k += 1
    print("encode soak (cpu): %d clips, mismatches %d, %.0f s" % (n, bad, time.time() - t0))
    sys.exit(1 if bad else 0)
if len(sys.argv) > 2 and sys.argv[2] == "rom":
    from oracle import ref
    import rombuild as rb, romscen, simutil
    budget = float(sys.argv[1])
    t0=time.time(); bad=0; n=0; k=0
    while time.time()-t0 < budget:
        osv=[rb.OS94, rb.OS95, rb.OS93B, rb.OS93A][k%4]
        seed=20000+k
        sc=romscen.make_scenario(os_version=osv, seed=seed, n_frames=400, with_errors=(k%3==0), version=(0x0105 if (osv==rb.OS95 and k%8==1) else None))
        rp=ref.RomPlayer(sc["images"], sc["master_volume"])
        want=rp.render_timeline(sc["writes"], sc["n_frames"]); hbw=rp.host_bytes(); rp.close()
        pcm,res,info,hb=simutil.rom_render(sc["images"], [(sc["writes"], sc["n_frames"], sc["master_volume"])])
        if not (np.array_equal(pcm[0],want) and hb==hbw):
            bad+=1; print("MISMATCH", hex(osv), seed, flush=True)
        n+=1; k+=1
    print("rom soak: %d scenarios, mismatches %d, %.0f s"%(n,bad,time.time()-t0))
    sys.exit(0)
t0=time.time(); n=0; bad=0; seed=10000
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 1500.0
while time.time()-t0 < budget:
    rng=np.random.default_rng(seed)
    streams=[]
    for os_,d,label in dcsfuzz.corpus(seed, n_each=2, nframes=int(rng.integers(1,120))):
        streams.append((d, os_, int(rng.integers(0,256)), int(rng.integers(0,256)), int(rng.integers(0,4))))
    # truncate / corrupt a few
    for k in range(3):
        i=int(rng.integers(0,len(streams))); d,os_,v,l,t=streams[i]
        if k==0 and len(d)>40: d=d[:int(rng.integers(20,len(d)))]
        elif k==1 and len(d)>40:
            b=bytearray(d); b[int(rng.integers(18,len(d)))]^=1<<int(rng.integers(0,8)); d=bytes(b)
        streams[i]=(d,os_,v,l,t)
    pcm,offs,res,bp,bt=simutil.decode_streams(streams)
    for i,(d,os_,vol,lvl,tail) in enumerate(streams):
        nf=(d[0]<<8)|d[1]
        want,rc=orc.decode(d,os_,vol,lvl,nf+tail)
        got=pcm[offs[i]:offs[i]+want.size]
        st=res[i]["status"]
        if st in (0,-5) and not np.array_equal(got,want):
            bad+=1; print("MISMATCH seed",seed,"stream",i,hex(os_),"status",st,flush=True)
        if st==0:
            w2=ref.decode(d+bytes(8),os_,vol,lvl,nf+tail)
            if not np.array_equal(got,w2):
                bad+=1; print("REF MISMATCH seed",seed,"stream",i,hex(os_),flush=True)
        n+=1
    seed+=1
print("soak: %d streams, %d seeds, mismatches %d, %.0f s"%(n,seed-10000,bad,time.time()-t0))
