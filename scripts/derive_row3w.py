"""Derivation of the third-order W-method "ROW3w" used by csrc/solver.cu (tc_row3_*) and oracle/row3_oracle.py.

Unknowns: 4 stages, stage 4 re-uses the right-hand side of stage 3 (3 evaluations), gamma = 0.435866521508459.
Constraints: the eight order conditions of a W-method of order 3 (Hairer & Wanner, Solving ODE II, IV.7: they must hold
for ANY matrix in place of the Jacobian; a 3-stage search finds no solution) and R(inf) = 0 for the exact Jacobian.
The remaining freedom minimises the four order-4 error terms of the exact-Jacobian case plus the coefficient size
(SLSQP from random starts); only candidates with |R(z)| <= 1 on the imaginary and the negative real axis are kept.
The script prints the coefficients in the original form (b, alpha_ij, gamma_ij), checks the order numerically on a
non-autonomous Van der Pol problem with a wrong (scaled / zero) Jacobian, and writes /tmp/exp/w34.npz.  The
implementation form (at_ij, c_ij, m_i, e_i) quoted in the sources follows from Gamma^-1 (tests/test_host_logic.py
re-derives the order conditions from those constants)."""
import numpy as np
from scipy.optimize import least_squares, minimize
gam=0.435866521508459
s=4
def unpack(x):
    b=x[:4]; A=np.zeros((4,4)); G=np.zeros((4,4))
    A[1,0]=x[4]; A[2,0]=x[5]; A[2,1]=x[6]; A[3,:]=A[2,:]          # stage 4 re-uses F3
    G[1,0]=x[7]; G[2,0]=x[8]; G[2,1]=x[9]; G[3,0]=x[10]; G[3,1]=x[11]; G[3,2]=x[12]
    G+=gam*np.eye(4)
    return b,A,G
def cons(x):
    b,A,G=unpack(x); al=A.sum(1); gp=G.sum(1); B=A+G
    r=[b.sum()-1, b@al-0.5, b@gp, b@al**2-1/3, b@(A@al)-1/6, b@(A@gp), b@(G@al), b@(G@gp),
       1-b@np.linalg.solve(B,np.ones(4))]
    return np.array(r)
def err4(x):
    # order-4 conditions of a Rosenbrock method with exact J (beta = alpha+gamma)
    b,A,G=unpack(x); al=A.sum(1); B=A+G; be=B.sum(1)
    return np.array([b@al**3-1/4, b@(al*(A@be))-1/8, b@(A@al**2)-1/12, b@(B@(B@be))-1/24])
def obj(x): return 0.05*np.sum(x**2)+np.sum(err4(x)**2)*50
rng=np.random.default_rng(3); best=None
for trial in range(200):
    x0=rng.uniform(-1,1,13)
    try:
        r=minimize(obj,x0,constraints=[{'type':'eq','fun':cons}],method='SLSQP',options={'ftol':1e-15,'maxiter':800})
    except Exception: continue
    if np.linalg.norm(cons(r.x))<1e-10:
        rr=least_squares(lambda z: np.concatenate([cons(z), 1e-3*(z-r.x)]), r.x, xtol=1e-15,ftol=1e-15,gtol=1e-15).x
        if np.linalg.norm(cons(rr))<1e-13:
            b,A,G=unpack(rr); B=A+G
            # A-stability check on imaginary axis and negative real axis
            ok=True; mx=0
            for z in list(1j*np.logspace(-2,4,200))+list(-np.logspace(-2,6,200)):
                R=1+z*b@np.linalg.solve(np.eye(4)-z*B,np.ones(4)); mx=max(mx,abs(R))
            val=obj(rr)
            if mx<=1+1e-9 and (best is None or val<best[0]): best=(val,rr,mx)
print(best[0], best[2]); x=best[1]
print(repr(x)); print("cons",np.linalg.norm(cons(x)),"err4",err4(x))
b,A,G=unpack(x); al=A.sum(1); gp=G.sum(1)
M=np.array([np.ones(4), al, gp]); 
# embedded: min-norm solution of the 3 order-2 conditions different from b
bh=np.linalg.lstsq(M, np.array([1,0.5,0]), rcond=None)[0]
print("b",repr(b)); print("bh",repr(bh), np.linalg.norm(b-bh))
import os; os.makedirs('/tmp/exp', exist_ok=True); np.savez('/tmp/exp/w34.npz', x=x, bh=bh, gam=gam)
def f(t,y): return np.array([y[1], 5*(1-y[0]**2)*y[1]-y[0]+np.sin(t)])
def J(t,y): return np.array([[0,1],[-10*y[0]*y[1]-1, 5*(1-y[0]**2)]])
def step(t,y,h,scale=1.0):
    Jm=scale*J(t,y); W=np.eye(2)-h*gam*Jm; k=[]; F=[None]*4
    for i in range(4):
        if i==3: Fi=F[2]
        else:
            yi=y+h*sum((A[i,j]*k[j] for j in range(i)), np.zeros(2)); Fi=f(t+al[i]*h,yi)
        F[i]=Fi
        rhs=Fi+h*Jm@sum(((G[i,j])*k[j] for j in range(i)), np.zeros(2))
        k.append(np.linalg.solve(W,rhs))
    yn=y+h*sum(b[i]*k[i] for i in range(4)); ye=y+h*sum(bh[i]*k[i] for i in range(4))
    return yn, yn-ye
def run(h,scale):
    t=0.0;y=np.array([2.0,0.0]);n=int(round(1.0/h))
    for _ in range(n): y,e=step(t,y,h,scale); t+=h
    return y
from scipy.integrate import solve_ivp
ref=solve_ivp(f,[0,1],[2.0,0.0],rtol=1e-13,atol=1e-14,method='Radau',jac=J).y[:,-1]
for sc in (1.0,0.7,0.0,1.5):
    errs=[np.linalg.norm(run(h,sc)-ref) for h in (0.02,0.01,0.005,0.0025)]
    print("J scale",sc, ["%.2e"%e for e in errs], [round(float(np.log2(errs[i]/errs[i+1])),2) for i in range(3)])
y=np.array([2.0,0.0])
e=[np.linalg.norm(step(0.0,y,h,0.7)[1]) for h in (0.02,0.01,0.005)]; print("est", e, [np.log2(e[i]/e[i+1]) for i in range(2)])
