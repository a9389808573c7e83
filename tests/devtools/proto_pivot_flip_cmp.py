import re,sys
def parse(f):
    out=[]
    for l in open(f):
        m=re.search(r"r(\d+) b\s*(\d+) device\s+(\d+)/\s*(\d+) obj (\S+) \| proto (\S+)\s+it\s+(\d+) fact\s+(\d+) mods\s+(\d+) obj (\S+) kkt (\S+)",l)
        if m: out.append(m.groups())
    return out
files=sys.argv[1:]
P=[parse(f) for f in files]
for k in range(len(P[0])):
    g=P[0][k]
    s=f"r{g[0]:>2s} b{g[1]:>4s} dev {g[2]:>3s}/{g[3]:>3s} {float(g[4]):+.5e}"
    for p in P:
        if k < len(p):
            q=p[k]; s+=f" | {q[5][:6]:6s} {q[6]:>3s}/{q[7]:>3s} m{q[8]:>4s} {float(q[9]):+.5e}"
    print(s)
for f,p in zip(files,P):
    print(f, "iters", sum(int(q[6]) for q in p), "facts", sum(int(q[7]) for q in p), "slow8x3 iters", sum(int(q[6]) for q in p if int(q[2])>45), "max", max(int(q[6]) for q in p), "nonsolved", sum(q[5]!='SOLVED' for q in p))
