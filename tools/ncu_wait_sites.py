import subprocess,csv,io,sys
rep=sys.argv[1]; which=int(sys.argv[2]) if len(sys.argv)>2 else 0; thr=float(sys.argv[3]) if len(sys.argv)>3 else 0.8
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
his=[i for i,r in enumerate(rows) if "Instructions Executed" in r]
hi=his[which]; end=his[which+1]-1 if which+1<len(his) else len(rows)
print(rows[hi-1][:2] if hi>0 else '')
h=rows[hi]; ie=h.index("Instructions Executed"); isrc=h.index("Source"); isamp=h.index("# Samples")
ins=[(r[isrc].strip(),int(r[ie]),int(r[isamp])) for r in rows[hi+1:end] if len(r)>ie and r[isamp].isdigit()]
tot=sum(s for _,_,s in ins); print('samples',tot,'warp-inst',sum(e for _,e,_ in ins))
for i,(src,e,s) in enumerate(ins):
    op=src.split()[0] if not src.startswith('@') else src.split()[1]
    if op.startswith(('LDTM','BAR.SYNC','UTCBAR','UBLKCP')) or (op.startswith('SYNCS.PHASECHK') and e>0) or s>thr/100*tot:
        print(f"{i:5d} {100*s/tot:5.1f}% exec {e:9d}  {src[:90]}")
