import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import support as S, kmersgwas_b200 as kg
np.set_printoptions(linewidth=250)
n_file=int(sys.argv[1]) if len(sys.argv)>1 else 64
n_rows=int(sys.argv[2]) if len(sys.argv)>2 else 128
table=S.synth_table(3,n_rows,n_file)
idx=np.arange(n_file)
K_o,cnt_o=S.oracle_kinship(table,n_file,idx//64,idx%64,1)
bits=np.unpackbits(np.ascontiguousarray(table[:,1:]).view(np.uint8),axis=1,bitorder='little')[:,:n_file].astype(np.int64)
keep=(bits.sum(1)>=1)&(bits.sum(1)<=n_file-1)
G=bits[keep].T@bits[keep]
ctx=kg.Context.identity(n_file)
ctx.set_option(kg.OPT_KINSHIP_ENGINE,2)
import torch
acc=torch.zeros(ctx.kinship_accum_len(),dtype=torch.int64,device='cuda')
ctx.kinship_begin(1,acc.data_ptr())
ctx.kinship_submit(table,n_rows); ctx.sync()
a=acc.cpu().numpy()
Gd=a[:-1].reshape(n_file,n_file)
print('kept',a[-1],keep.sum())
print('G ref[:10,:10]\n',np.tril(G)[:10,:10]); print('G dev[:10,:10]\n',Gd[:10,:10])
d=(Gd!=np.tril(G)); print('mismatch count',d.sum(),'of',d.size)
print('diag ref',np.diag(G)[:16]); print('diag dev',np.diag(Gd)[:16])
