import json,sys
d=json.load(open(sys.argv[1]))
import signal; signal.signal(signal.SIGPIPE, signal.SIG_DFL)
print("fps",round(d["fps"],1),"total_ms",round(d["total_ms"],1))
for r in d["layers"]:
    if r["type"]=="conv": print(r["layer"],r["shape"],r["k"],r["ms"],round(r["steps_per_s"]/1e12,2))
