import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from bench import build_workload
from hic_genome_assembler_b200._lib import Handle
from hic_genome_assembler_b200 import part1, part2
wl=build_workload(16000,20,7,torch.device('cuda',0))
h=Handle(0); h.p1_upload_device(wl['A'].data_ptr(),16000)
res=part1.run_part1_resident(h,min_size=5,min_frac=0.0,psig=.05)
groups=part1.chromosome_groups_compact(res['cuts'],res['leaves'],wl['bin_ids'],wl['scaffold_code'],wl['scaffold_names'],wl['scaffold_sizes'])
h.p2_use_p1_matrix()
got,cost,rounds,stats=part2.order_groups_device(h,wl['row_of_bin_id'],groups,6,5)
print('rounds',list(rounds)); print('improving scan windows per group',list(stats[:,0]))
