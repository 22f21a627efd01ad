import random
class Bar:
    def __init__(s,count): s.count=count; s.pending=count; s.phase=0
    def arrive(s):
        s.pending-=1
        assert s.pending>=0
        if s.pending==0: s.phase+=1; s.pending=s.count
    def test(s,parity):  # try_wait.parity: true if phase with this parity has completed
        return (s.phase & 1) != parity
NH=9
def consumer(b, npairs, nvalid_last=2):
    par_t=[0,0]; par_acc=[[0,0],[0,0]]
    for n in range(npairs):
        nv = nvalid_last if n==npairs-1 else 2
        for s in range(nv):
            yield ('wait', b['t_full'][s], par_t[s], f'C tfull n{n} s{s}'); par_t[s]^=1
            yield ('arrive', b['t_empty'][s])
            yield ('arrive', b['a_ready'][s])
        for e in range(1,NH):
            for s in range(nv):
                yield ('wait', b['acc'][s][0], par_acc[s][0], f'C acclo n{n} e{e} s{s}')
                # c0 compute, wait hi
                yield ('wait', b['acc'][s][1], par_acc[s][1], f'C acchi n{n} e{e} s{s}')
                yield ('arrive', b['d_free'][0])   # c==1
                yield ('step',)
                yield ('arrive', b['d_free'][1])   # c==3
                par_acc[s][0]^=1; par_acc[s][1]^=1
                yield ('arrive', b['a_ready'][s])
        for s in range(nv):
            yield ('wait', b['acc'][s][0], par_acc[s][0], f'C head lo n{n} s{s}')
            yield ('arrive', b['d_free'][0])
            yield ('wait', b['acc'][s][1], par_acc[s][1], f'C head hi n{n} s{s}')
            yield ('arrive', b['d_free'][1])
            par_acc[s][0]^=1; par_acc[s][1]^=1
def issuer(b, npairs, nvalid_last=2):
    par_a=[0,0]; par_free=[1,1]
    for n in range(npairs):
        nv = nvalid_last if n==npairs-1 else 2
        for e in range(1,NH+1):
            for s in range(nv):
                yield ('wait', b['a_ready'][s], par_a[s], f'I aready n{n} e{e} s{s}'); par_a[s]^=1
                for h in range(2):
                    yield ('wait', b['d_free'][h], par_free[h], f'I dfree n{n} e{e} s{s} h{h}'); par_free[h]^=1
                    yield ('commit', b['acc'][s][h])
def producer(b, npairs, nvalid_last=2):
    par_e=[1,1]
    for n in range(npairs):
        nv = nvalid_last if n==npairs-1 else 2
        for s in range(nv):
            yield ('wait', b['t_empty'][s], par_e[s], f'P tempty n{n} s{s}'); par_e[s]^=1
            yield ('arrive', b['t_full'][s])
def run(seed, npairs=3, nvalid_last=2, commit_delay=True):
    rnd=random.Random(seed)
    b={'t_full':[Bar(1),Bar(1)],'t_empty':[Bar(4),Bar(4)],'a_ready':[Bar(4),Bar(4)],'d_free':[Bar(4),Bar(4)],'acc':[[Bar(1),Bar(1)],[Bar(1),Bar(1)]]}
    procs=[consumer(b,npairs,nvalid_last) for _ in range(4)]+[issuer(b,npairs,nvalid_last),producer(b,npairs,nvalid_last)]
    cur=[next(p) for p in procs]
    pending_commits=[]  # in-order async commits
    done=[False]*len(procs)
    idle=0
    while not all(done):
        # async commits complete in order at random times
        if pending_commits and rnd.random()<0.3:
            pending_commits.pop(0).arrive()
        i=rnd.randrange(len(procs))
        if done[i]: continue
        op=cur[i]
        prog=False
        if op[0]=='wait':
            if op[1].test(op[2]): prog=True
        elif op[0]=='arrive': op[1].arrive(); prog=True
        elif op[0]=='commit': pending_commits.append(op[1]); prog=True
        else: prog=True
        if prog:
            idle=0
            try: cur[i]=next(procs[i])
            except StopIteration: done[i]=True
        else:
            idle+=1
            if idle>20000 and not pending_commits:
                return [c[3] if c[0]=='wait' else c for c,d in zip(cur,done) if not d]
    return None
bad=0
for seed in range(3000):
    r=run(seed, npairs=rnd if False else 1+seed%3, nvalid_last=1+(seed//3)%2)
    if r: bad+=1; print(seed, r)
print('bad',bad)
