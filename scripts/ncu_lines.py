"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per source line."""
import csv, sys
def num(x):
    try: return int(x)
    except: return 0
rows=list(csv.reader(open(sys.argv[1])))
top=int(sys.argv[2]) if len(sys.argv)>2 else 40
cur=None; agg=[]; hdr=None
for r in rows:
    if r and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r and r[0]=='Line No': hdr=r; continue
    if r and hdr and len(r)>7 and r[0].isdigit():
        agg.append((cur,int(r[0]),r[1][:100],num(r[4]),num(r[7])))
# inst executed on line rows is 0: sum from sass rows
cur=None; line=None; inst={}
for r in rows:
    if r and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r and r[0]=='Line No': continue
    if r and r[0].isdigit(): line=(cur,int(r[0])); continue
    if r and len(r)>7 and r[0]=='' and r[2].startswith('0x'):
        inst[line]=inst.get(line,0)+num(r[7])
tot_s=sum(a[3] for a in agg) or 1; tot_i=sum(inst.values()) or 1
print('total samples',tot_s,'total warp inst',tot_i)
for a in sorted(agg,key=lambda x:-x[3])[:top]:
    print(f"{a[0]}:{a[1]:4d} s={a[3]/tot_s*100:5.1f}% i={inst.get((a[0],a[1]),0)/tot_i*100:5.1f}%  {a[2]}")
