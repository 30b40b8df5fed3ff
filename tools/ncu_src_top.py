import csv, sys, collections
csv.field_size_limit(10**9)
want_kernel = sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 25
cur_file=None; cur_fn=None; hdr=None
agg=collections.defaultdict(lambda:[0,0,0,0,0,""])  # samples, inst, thr_inst, long_sb, short_sb
with open(sys.argv[3] if len(sys.argv) > 3 else '/tmp/p1_src_all.csv') as f:
    for r in csv.reader(f):
        if not r: continue
        if r[0]=="File Path": cur_file=r[1].split('/')[-1]; continue
        if r[0]=="Function Name": cur_fn=r[1]; continue
        if r[0]=="Line No": hdr=r; ci={n:hdr.index(n) for n in ("# Samples","Instructions Executed","Thread Instructions Executed","stall_long_sb","stall_short_sb","stall_lg","stall_mio","stall_wait","stall_not_selected","stall_math","stall_branch_resolving")}; continue
        if hdr is None or want_kernel not in (cur_fn or ""): continue
        if not r[0].strip().isdigit(): continue
        def num(x):
            try: return float(x)
            except: return 0.0
        k=(cur_file,int(r[0]))
        a=agg[k]
        a[0]+=num(r[ci["# Samples"]]); a[1]+=num(r[ci["Instructions Executed"]]); a[2]+=num(r[ci["Thread Instructions Executed"]])
        a[3]+=num(r[ci["stall_long_sb"]]); a[4]+=num(r[ci["stall_short_sb"]])+num(r[ci["stall_mio"]])
        if not a[5]: a[5]=r[1].strip()[:90]
ts=sum(a[0] for a in agg.values()) or 1; ti=sum(a[1] for a in agg.values()) or 1
print(f"kernel {want_kernel}: samples {ts:.0f} warp-inst {ti:.3g}")
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][0])[:top]:
    print(f"{100*a[0]/ts:5.1f}%smp {100*a[1]/ti:5.1f}%ins lsb {100*a[3]/ts:4.1f} ssb {100*a[4]/ts:4.1f} thr {a[2]/max(a[1],1):4.1f} {k[0]}:{k[1]} {a[5]}")
