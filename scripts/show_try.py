import json,sys
for f in sys.argv[1:]:
    print("==",f)
    try:
        d=json.load(open(f))
    except Exception as e:
        print("ERR",e); continue
    for r in d:
        print('%4s %6d %3d %3d %5s tc %6.1f alt %6.1f us  x%.2f  GB/s %4.0f  err tc %.1e alt %.1e  d %.1e dT %.1e'%(r['level'],r['n_out'],r['c_in'],r['c_out'],r['alt'],r['us_tc'],r['us_alt'],r['us_tc']/r['us_alt'],r['gbps_alt'],r.get('err_tc_vs_fp64',-1),r.get('err_alt_vs_fp64',-1),r['max_tc_minus_alt'],r['max_tc_minus_alt_mirrorT']))
