import sys, time, importlib
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import oracle_lib as O
rtb=O.rtb
t=time.time(); s=rtb.host.make_mesh_scene(max_bvh_depth=32, subdivisions=6); print('build', time.time()-t, len(s.triangles), len(s.nodes))
W,H,spp=128,72,8
p=rtb.host.make_params(s,W,H,spp,50,aperture=0.05)
ref=O.Buffers(W,H); t=time.time(); O.sample_batch(s,p,ref); print('oracle', time.time()-t)
ctx=rtb.plugin.Context(0); ctx.upload(s)
for k in (rtb.abi.KERNEL_SIMPLE, rtb.abi.KERNEL_MEGA):
    ctx.set_option(rtb.abi.OPT_KERNEL,k)
    got=rtb.plugin.HostBuffers(W,H); ctx.sample_batch(p,got)
    print(k, np.array_equal(ref.out_color[:,3],got.out_color[:,3]), np.array_equal(ref.diagnostics['ray_count'],got.diagnostics['ray_count']), float(np.abs(ref.rgb()-got.rgb()).max()), ctx.last_kernel_ms())
W,H,spp=1920,1080,64
p=rtb.host.make_params(s,W,H,spp,50,aperture=0.05)
ctx.set_option(rtb.abi.OPT_KERNEL,rtb.abi.KERNEL_MEGA)
b=rtb.plugin.HostBuffers(W,H)
for _ in range(2):
    ctx.sample_batch(p,b); print('1080p x64', ctx.last_kernel_ms(), W*H*spp/ctx.last_kernel_ms()/1e3,'Msamples/s')
