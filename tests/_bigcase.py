import sys, os, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import parity_util as pu
pu.CASES['big'] = (dict(total_len=5_000_000, n_contigs=2, seed=21, repeat_frac=0.02),
                   dict(n=100000, read_len=150, paired=True, seed=23, sub=0.01, indel=0.001), ["-K", "10000000"])
t=time.time(); c = pu.make_case('/tmp/b2cases', 'big'); print('gen', time.time()-t)
import subprocess
t=time.time(); ref = pu.oracle_sam(c, extra=["-t", "16"]); print('oracle -t16', time.time()-t)
t=time.time()
r = subprocess.run([pu.CLI, "mem"] + c["opts"] + ["-f", c["dir"]+"/engine.sam", c["ref"]] + c["fq"], stderr=subprocess.PIPE, env=dict(os.environ, B200BWA_TIMES="1"))
print('engine', time.time()-t, 'rc', r.returncode)
print("\n".join(l for l in r.stderr.decode().splitlines() if l.startswith('[T::') or 'Real time' in l or 'E::' in l))
d = pu.first_diff(ref, c["dir"]+"/engine.sam")
print('IDENTICAL' if not d else d)
