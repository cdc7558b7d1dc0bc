import json,sys
for f in sys.argv[1:]:
    txt=[l for l in open(f) if l.startswith("{")][-1]
    d=json.loads(txt)
    st=d["roofline_stages"]
    print(f, "N=%d value=%.2f ms=%.2f e2e=%.2f"%(d["n_gpus"],d["value"],d["ms_per_step"],d["e2e"]["value"]), " | ".join("%s %.2fx%d"%(k.replace("plane_moments_","pm_"),v["ms"],v["launches_per_step"]) for k,v in st.items()), "nvlink", st.get("a2a_pack",{}).get("nvlink_gbs_per_gpu"))
    if d.get("timeline"): print("   timeline", {k: ([round(x,2) for x in v] if isinstance(v, list) else (round(v,2) if isinstance(v,(int,float)) else v)) for k,v in d["timeline"].items()})
