"""Reads the byte code of a CPython 3.10 .pyc with any newer interpreter (TEST INFRASTRUCTURE, oracle/).

The reference's models/longformer_noffn.py has no source in the repository, only
models/__pycache__/longformer_noffn.cpython-310.pyc.  `marshal` of Python 3.12 cannot load 3.10 code objects, so this
file carries its own reader of the 3.10 marshal format and the 3.10 opcode table and prints a disassembly; the
restatement in oracle/ref_torch.NoffnLayer and multimodaltopicsegmentation_b200/recurrent_longformer.py follows it:

    python oracle/dis_pyc310.py /root/reference/models/__pycache__/longformer_noffn.cpython-310.pyc out.txt
    (look at LongformerSelfAttention.forward, LongformerAttention.__init__/forward, LongformerLayer.__init__/forward)
"""
import struct, sys
OP = {1:'POP_TOP',2:'ROT_TWO',3:'ROT_THREE',4:'DUP_TOP',5:'DUP_TOP_TWO',6:'ROT_FOUR',9:'NOP',10:'UNARY_POSITIVE',11:'UNARY_NEGATIVE',12:'UNARY_NOT',15:'UNARY_INVERT',16:'BINARY_MATRIX_MULTIPLY',17:'INPLACE_MATRIX_MULTIPLY',19:'BINARY_POWER',20:'BINARY_MULTIPLY',22:'BINARY_MODULO',23:'BINARY_ADD',24:'BINARY_SUBTRACT',25:'BINARY_SUBSCR',26:'BINARY_FLOOR_DIVIDE',27:'BINARY_TRUE_DIVIDE',28:'INPLACE_FLOOR_DIVIDE',29:'INPLACE_TRUE_DIVIDE',30:'GET_LEN',49:'WITH_EXCEPT_START',55:'INPLACE_ADD',56:'INPLACE_SUBTRACT',57:'INPLACE_MULTIPLY',59:'INPLACE_MODULO',60:'STORE_SUBSCR',61:'DELETE_SUBSCR',62:'BINARY_LSHIFT',63:'BINARY_RSHIFT',64:'BINARY_AND',65:'BINARY_XOR',66:'BINARY_OR',67:'INPLACE_POWER',68:'GET_ITER',71:'LOAD_BUILD_CLASS',74:'LOAD_ASSERTION_ERROR',75:'INPLACE_LSHIFT',76:'INPLACE_RSHIFT',77:'INPLACE_AND',78:'INPLACE_XOR',79:'INPLACE_OR',82:'LIST_TO_TUPLE',83:'RETURN_VALUE',84:'IMPORT_STAR',85:'SETUP_ANNOTATIONS',86:'YIELD_VALUE',87:'POP_BLOCK',89:'POP_EXCEPT',90:'STORE_NAME',91:'DELETE_NAME',92:'UNPACK_SEQUENCE',93:'FOR_ITER',94:'UNPACK_EX',95:'STORE_ATTR',96:'DELETE_ATTR',97:'STORE_GLOBAL',98:'DELETE_GLOBAL',99:'ROT_N',100:'LOAD_CONST',101:'LOAD_NAME',102:'BUILD_TUPLE',103:'BUILD_LIST',104:'BUILD_SET',105:'BUILD_MAP',106:'LOAD_ATTR',107:'COMPARE_OP',108:'IMPORT_NAME',109:'IMPORT_FROM',110:'JUMP_FORWARD',111:'JUMP_IF_FALSE_OR_POP',112:'JUMP_IF_TRUE_OR_POP',113:'JUMP_ABSOLUTE',114:'POP_JUMP_IF_FALSE',115:'POP_JUMP_IF_TRUE',116:'LOAD_GLOBAL',117:'IS_OP',118:'CONTAINS_OP',119:'RERAISE',121:'JUMP_IF_NOT_EXC_MATCH',122:'SETUP_FINALLY',124:'LOAD_FAST',125:'STORE_FAST',126:'DELETE_FAST',129:'GEN_START',130:'RAISE_VARARGS',131:'CALL_FUNCTION',132:'MAKE_FUNCTION',133:'BUILD_SLICE',135:'LOAD_CLOSURE',136:'LOAD_DEREF',137:'STORE_DEREF',138:'DELETE_DEREF',141:'CALL_FUNCTION_KW',142:'CALL_FUNCTION_EX',143:'SETUP_WITH',144:'EXTENDED_ARG',145:'LIST_APPEND',146:'SET_ADD',147:'MAP_ADD',148:'LOAD_CLASSDEREF',155:'FORMAT_VALUE',156:'BUILD_CONST_KEY_MAP',157:'BUILD_STRING',160:'LOAD_METHOD',161:'CALL_METHOD',162:'LIST_EXTEND',163:'SET_UPDATE',164:'DICT_MERGE',165:'DICT_UPDATE'}
CMP=['<','<=','==','!=','>','>=']
class Code:
    pass
class R:
    def __init__(s,b): s.b=b; s.p=0; s.refs=[]
    def u8(s): v=s.b[s.p]; s.p+=1; return v
    def i32(s): v=struct.unpack_from('<i',s.b,s.p)[0]; s.p+=4; return v
    def rd(s,n): v=s.b[s.p:s.p+n]; s.p+=n; return v
    def obj(s):
        t=s.u8(); flag=t&0x80; t=chr(t&0x7f)
        idx=None
        def reg(v):
            if flag: s.refs.append(v)
            return v
        if t=='0': return None
        if t=='N': return None
        if t=='F': return False
        if t=='T': return True
        if t=='.': return Ellipsis
        if t=='S': return StopIteration
        if t=='i': return reg(s.i32())
        if t=='l':
            n=s.i32(); neg=n<0; n=abs(n); v=0
            for k in range(n): v|=struct.unpack_from('<H',s.b,s.p)[0]<<(15*k); s.p+=2
            return reg(-v if neg else v)
        if t=='g': v=struct.unpack_from('<d',s.b,s.p)[0]; s.p+=8; return reg(v)
        if t=='s': n=s.i32(); return reg(s.rd(n))
        if t in 'tu': n=s.i32(); return reg(s.rd(n).decode('utf8','surrogatepass'))
        if t in 'aA': n=s.i32(); return reg(s.rd(n).decode('latin1'))
        if t in 'zZ': n=s.u8(); return reg(s.rd(n).decode('latin1'))
        if t in '()':
            n=s.u8() if t==')' else s.i32()
            if flag: idx=len(s.refs); s.refs.append(None)
            v=tuple(s.obj() for _ in range(n))
            if flag: s.refs[idx]=v
            return v
        if t=='[':
            n=s.i32(); v=[]; reg(v)
            for _ in range(n): v.append(s.obj())
            return v
        if t in '<>':
            n=s.i32()
            if flag: idx=len(s.refs); s.refs.append(None)
            v=frozenset(s.obj() for _ in range(n))
            if flag: s.refs[idx]=v
            return v
        if t=='{':
            v={}; reg(v)
            while True:
                k=s.obj()
                if k is None and s.b[s.p-1]==ord('0'): break
                v[k]=s.obj()
            return v
        if t=='r': return s.refs[s.i32()]
        if t=='c':
            if flag: idx=len(s.refs); s.refs.append(None)
            c=Code()
            c.argcount=s.i32(); c.posonly=s.i32(); c.kwonly=s.i32(); c.nlocals=s.i32(); c.stacksize=s.i32(); c.flags=s.i32()
            c.code=s.obj(); c.consts=s.obj(); c.names=s.obj(); c.varnames=s.obj(); c.freevars=s.obj(); c.cellvars=s.obj()
            c.filename=s.obj(); c.name=s.obj(); c.firstlineno=s.i32(); c.lnotab=s.obj()
            if flag: s.refs[idx]=c
            return c
        raise ValueError(f'type {t!r} at {s.p}')
def dis(c, out, indent=''):
    out.append(f"{indent}=== code {c.name} args={c.varnames[:c.argcount+c.kwonly]} line {c.firstlineno} flags={c.flags:#x}")
    code=c.code; ext=0
    for i in range(0,len(code),2):
        op,arg=code[i],code[i+1]|ext
        ext=(arg<<8) if op==144 else 0
        if op==144: continue
        n=OP.get(op,f'OP{op}'); a=''
        if op in (100,): 
            v=c.consts[arg]; a=repr(v) if not isinstance(v,Code) else f'<code {v.name}>'
        elif op in (90,91,95,96,97,98,101,106,108,109,116,160): a=c.names[arg]
        elif op in (124,125,126): a=c.varnames[arg]
        elif op in (135,136,137,138,148): a=(c.cellvars+c.freevars)[arg]
        elif op==107: a=CMP[arg]
        elif op in (110,93,122,143): a=f'to {i+2+arg*2}'
        elif op in (111,112,113,114,115,121): a=f'to {arg*2}'
        else: a=str(arg) if op>=90 else ''
        out.append(f"{indent}{i:5d} {n:24s} {a}")
    for v in c.consts:
        if isinstance(v,Code): dis(v,out,indent+'  ')
b=open(sys.argv[1],'rb').read()
r=R(b[16:])
top=r.obj()
out=[]
dis(top,out)
open(sys.argv[2],'w').write('\n'.join(out))
print(len(out),'lines')
