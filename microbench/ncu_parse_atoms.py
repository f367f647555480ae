import sys,csv
rows=list(csv.reader(sys.stdin))
h=[r for r in rows if "Metric Name" in r]
if h:
    h=h[0]; ki=h.index("ID"); mi=h.index("Metric Name"); vi=h.index("Metric Value")
    d={}
    for r in rows:
        if len(r)==len(h) and r[ki].isdigit(): d.setdefault(int(r[ki]),{})[r[mi]]=float(r[vi].replace(",",""))
    for k in sorted(d):
        m=d[k]; a=m["smsp__inst_executed_op_shared_atom.sum"]; w=m["l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum"]
        p0=int(sys.argv[1]) if len(sys.argv)>1 else 0
        print("launch %2d pat %2d NA %d  wavefronts/instr %.3f  dur_us %.1f" % (k, p0+k//2, 1+k%2, w/a, m["gpu__time_duration.sum"]/1e3))
