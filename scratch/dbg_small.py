import numpy as np, sys
sys.path.insert(0,"tests")
from golden_util import Golden
from treepl_b200 import engine as eng
gd=Golden("M1155")
plp=eng.PLCalcParallel.from_flat(gd.flat)
fp,P=eng.generate_param_order_vector(gd.flat["free"],False)
x0=plp.set_freeparams(P,False,fp)
B=32
X=np.repeat(x0[:,None],B,axis=1); X[:116,:]=0.01
plp.batch_configure(B)
print("launching gen2 nomask", flush=True)
F,G=plp.calc_pl_grad_batch(X,mode=1); print(F[:3], plp.last_kernel_generation(), flush=True)
