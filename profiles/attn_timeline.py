"""Development aid: clock64() timeline of CTA (0,0) of the tensor-core attention kernel (pass B)."""
import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import context_transformer_b200 as ctx
from context_transformer_b200 import _lib
from oracle import synth
net = ctx.build_net(types.SimpleNamespace(method='ours', phase=2, setting='transfer', precision='bf16'), 300, 60)
net.load_state_dict(synth.seeded_state(net.state_dict(), seed=0)); net.eval(); net.device = 'cuda:0'; net.cuda(); net.use_cuda_graph = False
x = synth.seeded_input(32, 300, seed=0).cuda()
net(x); torch.cuda.synchronize()
buf = torch.zeros(1024, dtype=torch.int64, device='cuda')
_lib.lib().ctx_debug_set_attention_timeline(buf.data_ptr())
net(x); torch.cuda.synchronize()
_lib.lib().ctx_debug_set_attention_timeline(None)
ph = buf.cpu()[500:514]
print('phases (start, xload, xstore, xq_done, q_ready/passA, passB, passB end, epi weights, O loaded, end, z done, classifier done, softmax done, after barrier):', [int(x - ph[0]) for x in ph])
t = buf.cpu()[:512].view(32, 16)
names = ['mma:kv_full', 'mma:s_emptyA', 'mma:s_emptyB(QK_A issued)', 'mma:QK_B issued', 'mma:p_fullA', 'mma:p_fullB(PV_A issued)', 'mma:PV_B issued',
         'smA:top', 'smA:s_full', 'smA:ld0', 'smA:exp0', 'smA:pv_done', 'smA:ld1', 'smA:exp1', 'smA:st1']
t0 = int(t[4, 0])
for j in range(4, 9):
    print('j=%2d ' % j + ' | '.join('%s=%d' % (n, int(t[j, i]) - t0) for i, n in enumerate(names)))
