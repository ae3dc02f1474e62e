import sys, warnings; sys.path.insert(0,'/root/repo'); warnings.filterwarnings("ignore")
import torch
from oracle import fd_oracle as O
from pyapes_b200.geometry import Box
from pyapes_b200.mesh import Mesh
from pyapes_b200.solver.fdm import FDM
from pyapes_b200.solver.ops import Solver
from pyapes_b200.variables import Field
from pyapes_b200.variables.bcs import mixed_bcs
n=[40,36,48]; kinds=["dirichlet"]*6; vals=[0.0,0.25,0.0,0.0,-0.5,0.0]
mesh=Mesh(Box[0:1,0:2,0:1],None,n,"cuda","double")
var=Field("p",1,mesh,{"domain":mixed_bcs(vals,kinds),"obstacle":None})
g=torch.Generator().manual_seed(99); rhs_h=torch.rand(1,*n,generator=g,dtype=torch.float64)-0.5
s=Solver({"fdm":{"method":"bicgstab","tol":1e-8,"max_it":3000,"report":False}})
s.set_eq(FDM().laplacian(1.0,var)==rhs_h.to("cuda")); print("gpu",s.solve())
xs,dx=O.make_axes([0,0,0],[1,2,1],n); bcs=[O.FaceBC(f,k,v) for f,k,v in zip(O.FACES,kinds,vals)]
x0=torch.zeros(1,*n,dtype=torch.float64); eq=O.Equation([O.Term("laplacian",1.0,1.0)],dx,xs,bcs).build(x0)
rhs_o=eq.adjust_rhs(x0,rhs_h.clone()); sol,rep,_=O.bicgstab(eq,x0,rhs_o,1e-8,3000); print("oracle",rep, (var().cpu()-sol).abs().max().item())
