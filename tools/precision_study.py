"""CPU study of cheaper numeric recipes for the tensor-core layers (torch fp8 dtypes, exact accumulation),
against the fp64 oracle logits of tests/golden/cnn_golden.npz.  Result (N=128): fp16x3 6.9e-5 max |dsoftmax|;
third pass in FP8 with a separately scaled accumulator 7.6e-4; FP8 sharing the accumulator 3.8e-3;
third pass dropped 2.0e-2.  Only fp16x3 leaves margin under the 1e-3 budget."""
import numpy as np, torch, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn.functional as F
from oracle import alexnet, encoder_c
from svision_b200 import weights
torch.set_num_threads(8)
w = weights.synthetic_weights()
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'cnn_golden.npz'))
N = int(os.environ.get('N', 128))
rows = g['rows'][:N]; ref = torch.from_numpy(g['logits_fp64'][:N])
imgs = torch.from_numpy(encoder_c.encode_f32(rows)).double().permute(0,3,1,2).contiguous()
D = torch.float64
def split16(x):
    hi = x.to(torch.float16).to(D); lo = (x - hi).to(torch.float16).to(D); return hi, lo
def q8(x, kind, scale):
    dt = torch.float8_e4m3fn if kind=='e4m3' else torch.float8_e5m2
    return (x*scale).to(torch.float32).to(dt).to(D)/scale
def layer(op, a, wt, mode):
    # a, wt float64 'true' values (a already = hi+lo representable), op(a,w) linear
    a_hi, a_lo = split16(a); w_hi, w_lo = split16(wt)
    out = op(a_hi, w_hi) + op(a_hi, w_lo)
    if mode == 'fp16x3': out = out + op(a_lo, w_hi)
    elif mode == 'fp8': out = out + op(q8(a_lo,'e4m3',2.0**10), q8(w_hi,'e5m2',2.0**-10)*1.0)
    elif mode == 'fp8b': out = out + op(q8(a_lo,'e4m3',2.0**10), q8(w_hi,'e4m3',2.0**6))   # separate-accumulator variant (free scaling)
    elif mode == '2pass': pass
    return out
def run(mode):
    def conv(name, x, stride=1, pad=0, groups=1, m=mode):
        wt = torch.from_numpy(w[name+'/weights']).double().permute(3,2,0,1).contiguous()
        b = torch.from_numpy(w[name+'/biases']).double()
        op = lambda A,Wt: F.conv2d(A, Wt, None, stride=stride, padding=pad, groups=groups)
        return F.relu(layer(op, x, wt, m) + b[None,:,None,None])
    def fc(name, x, relu=True, m=mode):
        wt = torch.from_numpy(w[name+'/weights']).double(); b = torch.from_numpy(w[name+'/biases']).double()
        y = layer(lambda A,Wt: A@Wt, x, wt, m) + b
        return F.relu(y) if relu else y
    lrn = lambda x: F.local_response_norm(x, 5, alpha=1e-4, beta=0.75, k=1.0)
    x = F.conv2d(imgs, torch.from_numpy(w['conv1/weights']).double().permute(3,2,0,1), torch.from_numpy(w['conv1/biases']).double(), stride=4)
    x = lrn(F.max_pool2d(F.relu(x),3,2))                       # conv1 exact-ish (fused front end)
    x = lrn(F.max_pool2d(conv('conv2', x, pad=2, groups=2),3,2))
    x = conv('conv3', x, pad=1); x = conv('conv4', x, pad=1, groups=2); x = conv('conv5', x, pad=1, groups=2)
    x = F.max_pool2d(x,3,2).permute(0,2,3,1).reshape(N,-1)
    x = fc('fc6', x); x = fc('fc7', x)
    wt8 = torch.from_numpy(w['fc8/weights']).double(); b8 = torch.from_numpy(w['fc8/biases']).double()
    hi, lo = split16(x); logits = (hi+lo)@wt8 + b8
    return logits
for mode in ('fp16x3','fp8','fp8b','2pass'):
    t=time.time(); l = run(mode)
    dl = (l-ref).abs().max().item(); dp = (torch.softmax(l,1)-torch.softmax(ref,1)).abs().max().item()
    flips = int((l.argmax(1)!=ref.argmax(1)).sum())
    print(f'{mode:8s} max|dlogit| {dl:.3e}  max|dsoftmax| {dp:.3e}  label flips {flips}/{N}  ({time.time()-t:.0f}s)', flush=True)
