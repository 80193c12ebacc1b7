import json, sys
sys.path.insert(0, '/root/repo')
import torch, bench
print(json.dumps(bench.plugin_api_timings(torch.device('cuda', 0)), indent=1))
