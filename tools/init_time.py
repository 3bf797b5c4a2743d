import time,torch,sys
torch.cuda.init()
from zkp_ecdsa_b200.api import Engine
t=time.time(); e=Engine(0); torch.cuda.synchronize(); t1=time.time()-t
t=time.time(); p=e.generate_params_list(80); torch.cuda.synchronize(); t2=time.time()-t
print(sys.argv[1],"init s",round(t1,3),"params s",round(t2,3), [round(x/2**30,1) for x in torch.cuda.mem_get_info()])
