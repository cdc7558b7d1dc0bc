import json,sys
for f in sys.argv[1:]:
    txt=[l for l in open(f) if l.startswith("{")][-1]
    d=json.loads(txt)
    st=d["roofline_stages"]
    print(f, "N=%d value=%.2f ms=%.2f e2e=%.2f"%(d["n_gpus"],d["value"],d["ms_per_step"],d["e2e"]["value"]), " | ".join("%s %.2fx%d"%(k.replace("plane_moments_","pm_"),v["ms"],v["launches_per_step"]) for k,v in st.items()), "nvlink", st.get("a2a_pack",{}).get("nvlink_gbs_per_gpu"))
