import sys
lines=open(sys.argv[1]).read().split('=== timed launch ===')[1].strip().splitlines()
ev=[]
for l in lines:
    if l.startswith('TL'):
        d=dict(x.split('=') for x in l.split()[1:])
        ev.append((int(d['t']),int(d['role']),int(d['tag']),int(d['k'])))
ev.sort()
names={9:'E0 arrive',41:'loads issued',40:'ctx->acc0 done',30:'zwin done',100:'M0 issue-start',101:'M1 issue-start',102:'M2 issue-start',200:'M0 committed',201:'M1 committed',202:'M2 committed',31:'L zempty ok',32:'L done',10:'E0 enter',11:'E1 enter',12:'E2 enter',41:'E1 accfull ok(h)',50:'E0 hempty ok',51:'E1 ready',52:'E2 ready',20:'E0 done',21:'E1 done',22:'E2 done'}
for t,r,tag,k in ev:
    print("%7d  role%d  %-16s k=%d" % (t,r,names.get(tag,str(tag)),k))
