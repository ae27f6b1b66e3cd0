"""Per-source-line view of a kernel from an `ncu --set full --import-source on` report (kernels built with -lineinfo).

usage: tools/ncu_lines.py <report.ncu-rep> <source file> [top N]
Reads `ncu -i <rep> --page source --csv --print-source cuda,sass` for the kernels matching `em_kernel`, sums the warp
stall samples, executed instructions and barrier stalls of the SASS lines behind every CUDA source line and prints the
top N lines by samples (used for profiles/*_lines.txt).
"""
import csv, collections, sys, subprocess
rep, src_path = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv)>3 else 40
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","--kernel-name","regex:em_kernel"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur_file=None; cur_line=None; agg=collections.defaultdict(lambda:[0,0,0])
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur_file=r[1].split("/")[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; si=r.index("# Samples"); ie=r.index("Instructions Executed"); sb=r.index("stall_barrier"); continue
    if r[0]!="":
        cur_line=r[0]; continue
    try: s=int(r[si]); e=int(r[ie]); b=int(r[sb])
    except: continue
    k=(cur_file,int(cur_line)); agg[k][0]+=s; agg[k][1]+=e; agg[k][2]+=b
tots=sum(v[0] for v in agg.values()); tote=sum(v[1] for v in agg.values())
print("samples",tots,"instr %.1fM"%(tote/1e6))
src=open(src_path).read().splitlines()
for (f,l),v in sorted(agg.items(), key=lambda x:-x[1][0])[:topn]:
    print(f,l,"%.1f%%s"%(100*v[0]/tots),"%.1f%%i"%(100*v[1]/tote), "bar",v[2], (src[l-1].strip()[:80] if f=="em.cu" else ""))
