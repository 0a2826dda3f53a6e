import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f))
        print(f, {k:round(v,1) for k,v in d["phases_ms"].items()}, round(sum(d["phases_ms"].values()),1))
        print("   ", [(k[0][:24],k[1],round(k[2],2)) for k in d["kernels"][:18]])
    except Exception as e:
        print(f, "ERR", e)
