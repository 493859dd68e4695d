import json,sys
bad=0;n=0
for l in open(sys.argv[1]):
    r=json.loads(l)
    if r.get("status")!="done" and r.get("status")!="started":
        print("NOT DONE", r); bad+=1
    if r.get("status")=="done":
        n+=1
        if not r.get("ok",False):
            bad+=1; print("FAIL", r.get("case"), {k:v for k,v in r.items() if k in("f32","bf16","error")})
print(sys.argv[1], "cases", n, "bad", bad)
