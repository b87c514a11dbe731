import json,sys
for f in sys.argv[1:]:
    d=json.loads(open(f).read().strip().split('\n')[-1])
    s=d['sub']
    print(f, 'C2', d['roofline']['us_per_launch'], 'static', d['static_weights_pdl']['us_per_launch'],
          'C1', s['C1_gemv_M1_K4096_N4096_f16']['us_per_call'], 'C3', s['C3_gemv_M4_K4096_N4096_bias_bf16']['us_per_call'])
