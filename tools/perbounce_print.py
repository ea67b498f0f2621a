import csv, collections, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith("==")]
r=csv.DictReader(lines)
d=collections.OrderedDict()
for row in r:
    k=(row["ID"], row["Kernel Name"][:22])
    d.setdefault(k,{})[row["Metric Name"]]=row["Metric Value"]
short={"gpu__time_duration.sum":"ns","smsp__thread_inst_executed_per_inst_executed.ratio":"thr/inst","smsp__issue_active.avg.pct_of_peak_sustained_active":"issue%","sm__cycles_active.avg":"cyc_act","sm__cycles_elapsed.max":"cyc_el","smsp__inst_executed.sum":"inst","l1tex__t_sector_hit_rate.pct":"L1hit","sm__warps_active.avg.pct_of_peak_sustained_active":"warps%","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio":"longsb","lts__t_sector_hit_rate.pct":"L2hit"}
for k,v in d.items():
    print(k[0].rjust(3), k[1][5:20], "  ".join(f"{short[m]}={v[m]}" for m in short if m in v))
