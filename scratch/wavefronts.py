import numpy as np, sys
sys.path.insert(0,'.')
from fetalsyngen_b200.tables import make_affine_matrix
rs=np.random.RandomState(0)
S=256
layouts={'z32':(1,1,32),'z16y2':(1,2,16),'z8y4':(1,4,8),'z8y2x2':(2,2,8),'z4y4x2':(2,4,4),'z16x2':(2,1,16),'z4y8':(1,8,4), 'z2y4x4':(4,4,2)}
res={k:[] for k in layouts}
sect={k:[] for k in layouts}
for trial in range(200):
    rot=(2*20*rs.rand(3)-20)/180*np.pi
    A=make_affine_matrix(rot,0.04*rs.rand(3)-0.02,1+0.2*rs.rand(3)-0.1)
    for name,(lx,ly,lz) in layouts.items():
        # random tile origin
        o=rs.randint(40,200,size=3)
        g=np.stack(np.meshgrid(np.arange(lx),np.arange(ly),np.arange(lz),indexing='ij'),-1).reshape(-1,3)+o
        c=(A@(g-127.5).T).T+127.5
        f=np.floor(c).astype(int)
        lines=set(); 
        tot=0; tots=0
        for dx in (0,1):
            for dy in (0,1):
                for dz in (0,1):
                    p=f+np.array([dx,dy,dz])
                    addr=(p[:,0]*S+p[:,1])*S+p[:,2]
                    tot+=len(set(addr>>5)); tots+=len(set(addr>>3))
        res[name].append(tot/8); sect[name].append(tots/8)
for k in layouts: print(f"{k:8s} lines/request {np.mean(res[k]):5.2f}  sectors/request {np.mean(sect[k]):5.2f}")
