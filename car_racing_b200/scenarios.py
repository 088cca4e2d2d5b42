"""Seeded synthetic scenario generators for the BASELINE.json configs.

Host-side numpy only; used by bench.py and the tests so that every consumer draws the same
inputs.  Nothing here touches the GPU.
"""
import numpy as np

LAP_LENGTH = {"l_shape": 19.2296, "goggle": 19.1313, "ellipse": 25.4248, "m_shape": 49.8402}

# data/sys/LTI/matrix_{A,B}.csv of the reference: the identified LTI
# model every optimiser of the reference uses.  Kept as literals because /root/reference does
# not exist on the GPU box.
LTI_A = np.array([
    [9.671290682499817937e-01, 4.242128062809466527e-02, -1.334296593039446290e-02, -3.371777265649055985e-03, 3.227927831414771170e-06, 1.686059132152231384e-03],
    [-4.741648280908570059e-03, -2.468400046155550531e-01, 3.245228399640302103e-02, -6.761389137462380265e-04, 1.409084040214530391e-05, 5.146180819166828319e-03],
    [-4.687236221283337667e-02, -2.359785858136580039e+00, 3.057730658931238632e-01, -6.291019348164512102e-03, 1.407919154285720311e-04, 5.161261275626113226e-02],
    [1.491341619436537154e-02, -7.246004296693270286e-01, 9.814447747880826467e-02, 1.067946803205831907e+00, 3.685185147970903791e-06, 7.050901769705017648e-03],
    [9.697109942460338528e-02, -1.868558231814464871e-02, 1.741062754045796974e-03, -4.872250833025211156e-03, 1.000000846363945595e+00, -1.023352728807257542e-02],
    [6.714003570349933153e-04, -8.179829686470431460e-02, 1.271581430918338092e-02, 4.104114675570305626e-02, -1.886060580647413303e-06, 1.000653657321968648e+00],
])
LTI_B = np.array([
    [1.487289292077178041e-02, 9.770355093588267703e-02],
    [1.823272351013625059e-01, -9.009215991985008781e-04],
    [1.575741357076274385e+00, -9.783864169291941332e-03],
    [1.258236266910766066e-01, 7.018480442549679963e-04],
    [5.201009225593251689e-04, 4.856148476338860952e-03],
    [1.639395364567048513e-02, 1.979138102590245085e-07],
])


def default_cbf_params(N=20, **over):
    """Defaults of MPCCBFRacingParam / SystemParam / CarParam (base.py:272-291, 708-713, 699-705)
    and the constants hard-coded in control.mpccbf (control.py:527-528, 560)."""
    p = dict(A=LTI_A, B=LTI_B, Q=np.diag([10.0, 0.0, 0.0, 4.0, 0.0, 40.0]), R=np.diag([0.1, 0.1]), N=N,
             umax=[0.5, 1.0], vmin=0.0, vmax=10.0, width=1.0, alpha=0.8, margin=0.2, L=0.4, W=0.2, slack_w=1e4)
    p.update(over)
    return p


def _h0(ds, de, L=0.4, W=0.2, margin=0.2):
    return (ds / L) ** 6 + (de / W) ** 6 - 1.0 - margin


def mpccbf_scenarios(B, N=20, M=3, seed=1, track="l_shape", near_target_frac=0.25, vt=0.8):
    """Config 2: B random (x0, 3 static rivals) instances on l_shape, width 1.0.

    x0: vx~U(.4,1.5) vy~U(-.05,.05) wz~U(-.2,.2) epsi~U(-.1,.1) s~U(0,lap-5) ey~U(-.6,.6);
    rivals: s_j = s + U(0.6, 1.95 vx) kept inside the +-2vx proximity window (control.py:499-523),
    ey_j~U(-.7,.7), resampled while h(x0)<0.5 or |ey_j-ey|<0.15.  A fraction `near_target_frac`
    of the instances starts within +-10% of the target speed / centre line with rivals >=0.5 m
    to the side, so that a share of the optima is unsaturated (hard part 8).
    Returns x0 (B,6), xt (6,), obs (B,M,2,N+1), lap_off (B,M).
    """
    rng = np.random.default_rng(seed)
    lap = LAP_LENGTH[track]
    x0 = np.zeros((B, 6))
    obs = np.zeros((B, M, 2, N + 1))
    for b in range(B):
        near = rng.uniform() < near_target_frac
        while True:
            if near:
                vx = vt * rng.uniform(0.9, 1.1)
                x = np.array([vx, rng.uniform(-.01, .01), rng.uniform(-.02, .02), rng.uniform(-.01, .01),
                              rng.uniform(0, lap - 5), rng.uniform(-.02, .02)])
            else:
                vx = rng.uniform(0.4, 1.5)
                x = np.array([vx, rng.uniform(-.05, .05), rng.uniform(-.2, .2), rng.uniform(-.1, .1),
                              rng.uniform(0, lap - 5), rng.uniform(-.6, .6)])
            ok = True
            for j in range(M):
                for _ in range(200):
                    ds = min(rng.uniform(0.6, max(0.7, 1.95 * vx)), 1.98 * vx)
                    ey = rng.uniform(-0.7, 0.7)
                    far = abs(ey - x[5]) >= (0.5 if near else 0.15)
                    if _h0(ds, x[5] - ey) >= 0.5 and far:
                        break
                else:
                    ok = False
                obs[b, j, 0, :] = x[4] + ds
                obs[b, j, 1, :] = ey
            if ok:
                break
        x0[b] = x
    xt = np.array([vt, 0.0, 0.0, 0.0, 0.0, 0.0])
    return x0, xt, obs, np.zeros((B, M))


def ilqr_scenarios(B, N=50, seed=1, track="l_shape"):
    """Config 5: x0 as config 2, one static rival ahead s+U(0.6,1.6), ey_r~U(-.7,.7)."""
    rng = np.random.default_rng(seed)
    lap = LAP_LENGTH[track]
    x0 = np.zeros((B, 6))
    obs = np.zeros((B, 2, N + 1))
    for b in range(B):
        vx = rng.uniform(0.4, 1.5)
        x0[b] = [vx, rng.uniform(-.05, .05), rng.uniform(-.2, .2), rng.uniform(-.1, .1),
                 rng.uniform(0, lap - 5), rng.uniform(-.6, .6)]
        obs[b, 0, :] = x0[b, 4] + rng.uniform(0.6, 1.6)
        obs[b, 1, :] = rng.uniform(-0.7, 0.7)
    xt = np.array([0.8, 0.0, 0.0, 0.0, 0.0, 0.0])
    return x0, xt, obs, np.zeros(B)
