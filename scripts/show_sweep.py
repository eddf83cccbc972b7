import json, sys
d=json.load(open(sys.argv[1]))
print("HEAD")
for r in d['head']:
    print(f"B={r['B']} K={r['K']:2d} coh={int(r['coherent'])} tile={r['tcy']}x{r['tcx']} lpr={r['lpr']} {r['mode']:6s} call={r['call_ms']*1e3:7.1f}us kern={r['kernel_ms']*1e3:7.1f}us  {r['gpx_s']/1e3:6.1f} Gpx/s frac={r['frac']*100:5.2f}%")
print("HIST")
for r in d['hist']:
    print(f"{r['rows']}x{r['cols']} lut={int(r['lut'])} coh={str(r['coherent'])[:5]} mode={r['mode']} warps={r['warps']:2d} unroll={r['unroll']} kern={r['kernel_ms']*1e3:7.1f}us {r['gbs']:7.1f} GB/s frac={r['frac']*100:5.1f}%")
