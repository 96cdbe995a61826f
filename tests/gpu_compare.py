import sys, time
sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
import harness as H
from forge2d_b200 import scenes
name = sys.argv[1]; frames = int(sys.argv[2]); kw = eval(sys.argv[3]) if len(sys.argv) > 3 else {}
mode = int(sys.argv[4]) if len(sys.argv) > 4 else -1
every = int(sys.argv[5]) if len(sys.argv) > 5 else 1
ref = H.load('reference'); gpu = H.load('product')
print('has device', gpu.f2dHasDevice(), 'missing', gpu.missing)
sa = scenes.SCENES[name](ref, **kw); sb = scenes.SCENES[name](gpu, **kw)
gpu.f2dWorld_SetLaunchMode(sb.world, mode)
tr = tg = 0.0
ok = True
for f in range(frames):
    t0=time.perf_counter(); sa.step(); t1=time.perf_counter(); sb.step(); t2=time.perf_counter()
    tr += t1-t0; tg += t2-t1
    if f % every == 0 or f == frames-1:
        d = H.diff(H.snapshot(ref, sa.world), H.snapshot(gpu, sb.world))
        err = gpu.f2dGetLastError()
        if d or err:
            print(name, 'frame', f, 'DIFF', len(d), 'err', err)
            for x in d[:20]: print('   ', x)
            ok = False
            break
if ok: print(name, kw, 'mode', mode, 'all', frames, 'frames identical; mean ref %.3f ms gpu %.3f ms (incl. sync+header)' % (tr/frames*1e3, tg/frames*1e3))
