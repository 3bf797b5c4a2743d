// zkpAttestListGpu.ts — source-compatible TypeScript host over the N-API shim (zkattest_napi.cc).
//
// Drop next to src/zkpAttestList.ts of cloudflare/zkp-ecdsa v0.2.6 and re-export from src/index.ts: the four
// functions keep the reference's names and signatures (src/zkpAttestList.ts:88,94,104,147).  Result objects are
// rebuilt from the flat bytes with the reference's own deserializePoint / deserializeScalar, proofs handed to
// the verifier are serialised with the reference's own toBytes(), so writeJson / readJson / eq keep working.
// NOT compiled in this image (no node / tsc, SURVEY.md F7); the Python host zkp_ecdsa_b200/api.py mirrors it
// function for function and is what the tests exercise.
import { ExpProof } from './exp/exp.js'
import { GKProof } from './proofGK/gk.js'
import { PedersenParams } from './commit/pedersen.js'
import { PointAddProof } from './exp/pointAdd.js'
import { MultProof } from './commit/mult.js'
import { EqualityProof } from './commit/equality.js'
import { Group } from './curves/group.js'
import { SignatureProofList, SystemParametersList } from './zkpAttestList.js'
import { p256, tomEdwards256 } from './curves/instances.js'
import { posMod, rnd, toBytes } from './bignum/big.js'
// One native addon per ProofGroup: zkattest.node links libzkattest.so (tomEdwards256, what generateParamsList builds),
// zkattest_war256.node links libzkattest_war256.so (war256, the other group the JSON initialiser of
// curves/instances.ts:58-69 accepts for SystemParametersList.ProofGroup).  Same shim source, same exports; the
// war256 addon is loaded on first use.
// eslint-disable-next-line @typescript-eslint/no-require-imports
const native = require('../build/Release/zkattest.node')
let nativeWar: typeof native | undefined
function nativeFor(params: SystemParametersList): typeof native {
    const name = params.ProofGroup.c.name
    if (name === 'tomEdwards256') return native
    if (name !== 'war256') throw new Error(`invalid group name: ${name}`) // instances.ts:66
    // eslint-disable-next-line @typescript-eslint/no-require-imports
    if (nativeWar === undefined) nativeWar = require('../build/Release/zkattest_war256.node')
    return nativeWar
}

const NP = 65,
    NS = 32,
    V_SAMPLES = 20, // verifySignatureList's literal secparam (zkpAttestList.ts:177)
    IDX_PAD = 96

// ---- native parameter handles -------------------------------------------------------------------------------
// zka_params_create builds the fixed-base tables of NistGroup.h and ProofGroup.h in HBM: one native handle per
// parameter set.  At most MAX_HANDLES sets stay resident; the least recently used one is destroyed
// (zka_params_destroy) when a new one is needed.
const MAX_HANDLES = 4
const handles = new Map<string, { h: unknown; used: number; nat: typeof native }>()
let tick = 0
function paramsHandle(params: SystemParametersList): unknown {
    const hn = params.NistGroup.h.toBytes(),
        hp = params.ProofGroup.h.toBytes()
    const nat = nativeFor(params)
    const key = Buffer.from(hn).toString('hex') + Buffer.from(hp).toString('hex') + ':' + params.SecLevel
    let e = handles.get(key)
    if (e === undefined) {
        if (handles.size >= MAX_HANDLES) {
            let oldest: string | undefined, age = Infinity
            for (const [k, v] of handles) if (v.used < age) { age = v.used; oldest = k }
            if (oldest !== undefined) { const o = handles.get(oldest)!; o.nat.paramsDestroy(o.h); handles.delete(oldest) }
        }
        e = { h: nat.paramsCreate(hn, hp, params.SecLevel), used: 0, nat }
        handles.set(key, e)
    }
    e.used = ++tick
    return e.h
}

// ---- randomness tapes (include/zkattest.h) ---------------------------------------------------------------------
// rnd()'s rejection loop stays on the host (big.ts:171-181): the modulus of prover draw k depends only on k.
function proveTape(secLevel: number, n: number): Uint8Array {
    const draws = 3 + 4 * secLevel + 40 * secLevel + 5 * n,
        out = new Uint8Array(32 * draws)
    for (let k = 0; k < draws; k++) {
        const nist = k === 0 || (k >= 3 && k < 3 + 4 * secLevel && (k - 3) % 4 < 2)
        out.set(toBytes(rnd(nist ? p256.order : tomEdwards256.order), 32), 32 * k)
    }
    return out
}
// Verifier tape: 2n+1 Relation.drain scalars of verifyMembership (gk.ts:230,236,259), then the secLevel-2 index
// draws of generateIndices (exp.ts:101-106: rnd(limit - i), one byte each), padded to 96 bytes, then the
// 25 * 20 packed drain scalars of verifyExp.  The modulus of a packed drain depends on the challenge bits, which
// are not known yet: every 32-byte draw is taken below p256.n < tomEdwards256.order (2^-32 short of rnd()'s range).
function verifyTape(secLevel: number, n: number): Uint8Array {
    const g = 32 * (2 * n + 1),
        out = new Uint8Array(g + IDX_PAD + 32 * 25 * V_SAMPLES)
    for (let k = 0; k < 2 * n + 1; k++) out.set(toBytes(rnd(tomEdwards256.order), 32), 32 * k)
    for (let i = 0; i < secLevel - 2; i++) out[g + i] = Number(rnd(BigInt(secLevel - i)))
    for (let k = 0; k < 25 * V_SAMPLES; k++) out.set(toBytes(rnd(p256.order), 32), g + IDX_PAD + 32 * k)
    return out
}

// ---- flat proof layout (include/zkattest.h): reader ------------------------------------------------------------
class Reader {
    o = 0
    constructor(private b: Uint8Array, private pg: Group = tomEdwards256) {}
    take(n: number) {
        if (this.o + n > this.b.length) throw new Error('error deserializing Point')
        const v = this.b.subarray(this.o, this.o + n)
        this.o += n
        return v
    }
    np() {
        const v = this.take(NP)
        return v.every((x) => x === 0) ? p256.identity() : p256.deserializePoint(v)
    }
    wp() { return this.pg.deserializePoint(this.take(1 + 2 * this.pg.sizeFieldBytes())) } // 67 B tomEdwards256 / 65 B war256
    ns() { return p256.deserializeScalar(this.take(NS)) }
    ws() { return this.pg.deserializeScalar(this.take(this.pg.sizeFieldBytes())) } // 33 B / 32 B (group.ts:49-52)
}
function readMult(r: Reader) {
    const p = [r.wp(), r.wp(), r.wp(), r.wp(), r.wp(), r.wp()],
        s = [r.ws(), r.ws(), r.ws(), r.ws(), r.ws(), r.ws(), r.ws()]
    return new MultProof(p[0], p[1], p[2], p[3], p[4], p[5], s[0], s[1], s[2], s[3], s[4], s[5], s[6])
}
function readEq(r: Reader) { return new EqualityProof(r.wp(), r.wp(), r.ws(), r.ws(), r.ws()) }
export function readProof(bytes: Uint8Array, secLevel: number, pg: Group = tomEdwards256): SignatureProofList {
    const r = new Reader(bytes, pg),
        R = r.np(), comS1 = r.np(), kx = r.wp(), ky = r.wp(),
        exps: ExpProof[] = []
    for (let i = 0; i < secLevel; i++) {
        const tag = r.take(1)[0], A = r.np(), Tx = r.wp(), Ty = r.wp()
        if (tag === 1) exps.push(new ExpProof(A, Tx, Ty, r.ns(), r.ns(), r.ws(), r.ws()))
        else if (tag === 0) {
            const z = r.ns(), z2 = r.ns(),
                c = [r.wp(), r.wp(), r.wp(), r.wp()],
                m = [readMult(r), readMult(r), readMult(r), readMult(r)],
                ex = readEq(r), ey = readEq(r)
            exps.push(new ExpProof(A, Tx, Ty, undefined, undefined, undefined, undefined, z, z2,
                new PointAddProof(c[0], c[1], c[2], c[3], m[0], m[1], m[2], m[3], ex, ey), r.ws(), r.ws()))
        } else throw new Error('error deserializing Point')
    }
    const n = r.take(1)[0],
        pts = (k: number) => Array.from({ length: k }, () => r.wp()),
        scs = (k: number) => Array.from({ length: k }, () => r.ws())
    const gk = new GKProof(pts(n), pts(n), pts(n), pts(n), scs(n), scs(n), scs(n), r.ws())
    return new SignatureProofList(R, comS1, kx, ky, exps, gk)
}

// ---- flat proof layout: writer (the reference's own toBytes() in class-field order) ---------------------------
class Writer {
    parts: Uint8Array[] = []
    np(p: Group.Point) {
        const b = p.toBytes() // the P-256 identity is ONE 0x00 byte (weier.ts:247): a fixed slot holds 65 zero bytes
        this.parts.push(b.length === 1 ? new Uint8Array(NP) : b)
    }
    wp(p: Group.Point) { this.parts.push(p.toBytes()) }
    sc(s: Group.Scalar) { this.parts.push(s.toBytes()) } // 32 B (p256) / 33 B (tomEdwards256): group.ts:196-199
    byte(v: number) { this.parts.push(Uint8Array.of(v)) }
    mult(m: MultProof) {
        for (const p of [m.C_4, m.A_x, m.A_y, m.A_z, m.A_4_1, m.A_4_2]) this.wp(p)
        for (const s of [m.t_x, m.t_y, m.t_z, m.t_rx, m.t_ry, m.t_rz, m.t_r4]) this.sc(s)
    }
    eq(e: EqualityProof) { this.wp(e.A_1); this.wp(e.A_2); this.sc(e.t_x); this.sc(e.t_r1); this.sc(e.t_r2) }
    done(): Uint8Array {
        const n = this.parts.reduce((a, p) => a + p.length, 0), out = new Uint8Array(n)
        let o = 0
        for (const p of this.parts) { out.set(p, o); o += p.length }
        return out
    }
}
export function writeProof(proof: SignatureProofList): Uint8Array {
    const w = new Writer()
    w.np(proof.R); w.np(proof.comS1); w.wp(proof.keyXcom); w.wp(proof.keyYcom)
    for (const e of proof.expProof) {
        if (e.alpha && e.beta1 && e.beta2 && e.beta3) {
            w.byte(1); w.np(e.A); w.wp(e.Tx); w.wp(e.Ty)
            w.sc(e.alpha); w.sc(e.beta1); w.sc(e.beta2); w.sc(e.beta3)
        } else if (e.z && e.z2 && e.proof && e.r1 && e.r2) {
            w.byte(0); w.np(e.A); w.wp(e.Tx); w.wp(e.Ty)
            w.sc(e.z); w.sc(e.z2)
            const pa = e.proof
            for (const p of [pa.C_8, pa.C_10, pa.C_11, pa.C_13]) w.wp(p)
            for (const m of [pa.pi_8, pa.pi_10, pa.pi_11, pa.pi_13]) w.mult(m)
            w.eq(pa.pi_x); w.eq(pa.pi_y)
            w.sc(e.r1); w.sc(e.r2)
        } else throw new Error('params not found') // exp.ts:270,302
    }
    const gk = proof.membershipProof
    w.byte(gk.cl.length)
    for (const arr of [gk.cl, gk.ca, gk.cb, gk.cd]) for (const p of arr) w.wp(p)
    for (const arr of [gk.f, gk.za, gk.zb]) for (const s of arr) w.sc(s)
    w.sc(gk.zd)
    return w.done()
}

function ringBytes(keys: bigint[]): Uint8Array {
    const ring = new Uint8Array(32 * keys.length)
    keys.forEach((k, i) => ring.set(toBytes(posMod(k, tomEdwards256.order), 32), 32 * i)) // pad() wraps keys in newScalar (gk.ts:77)
    return ring
}

// ---- the reference's public functions ---------------------------------------------------------------------------
export function generateParamsList(secLevel = 80): SystemParametersList {
    const rnd64 = new Uint8Array(64)
    rnd64.set(toBytes(rnd(p256.order), 32), 0) // pedersen.ts:66 on p256, then on tomEdwards256 (zkpAttestList.ts:89-90)
    rnd64.set(toBytes(rnd(tomEdwards256.order), 32), 32)
    const { hNist, hProof } = native.paramsGenerate(rnd64)
    return new SystemParametersList(
        new PedersenParams(p256, p256.generator(), p256.deserializePoint(hNist)),
        new PedersenParams(tomEdwards256, tomEdwards256.generator(), tomEdwards256.deserializePoint(hProof)),
        secLevel)
}

export async function keyToInt(publicKey: CryptoKey): Promise<bigint> {
    const raw = new Uint8Array(await crypto.subtle.exportKey('raw', publicKey)) // zkpAttestList.ts:95
    const x: Uint8Array = await native.keyToInt(raw) // rejects with 'invalid public key'
    let v = 0n
    for (const b of x) v = (v << 8n) | BigInt(b)
    return v
}

export async function proveSignatureList(params: SystemParametersList, msgHash: Uint8Array, sigBytes: Uint8Array,
    publicKey: CryptoKey, which: number, keys: bigint[]): Promise<SignatureProofList> {
    const pk = new Uint8Array(await crypto.subtle.exportKey('raw', publicKey)) // zkpAttestList.ts:113
    const n = Math.ceil(Math.log2(keys.length)),
        res = await nativeFor(params).proveBatch(paramsHandle(params), msgHash, sigBytes, pk, Uint32Array.of(which), ringBytes(keys),
            proveTape(params.SecLevel, n), params.SecLevel)
    return readProof(res.proofs.subarray(0, res.lens[0]), params.SecLevel, params.ProofGroup.c)
}

export async function verifySignatureList(params: SystemParametersList, msgHash: Uint8Array, keys: bigint[],
    proof: SignatureProofList): Promise<boolean> {
    const n = Math.ceil(Math.log2(keys.length)),
        bytes = writeProof(proof),
        res = await nativeFor(params).verifyBatch(paramsHandle(params), msgHash, ringBytes(keys), bytes, Uint32Array.of(bytes.length),
            bytes.length, verifyTape(params.SecLevel, n), params.SecLevel)
    return res.ok[0] !== 0 // a status != 0 has already been rethrown as Error(<reference message>) by the shim
}

// Additive batch variants (B statements over one ring in one GPU pass).
export async function proveSignatureListBatch(params: SystemParametersList, msgHash: Uint8Array[], sigBytes: Uint8Array[],
    publicKey: CryptoKey[], which: number[], keys: bigint[]): Promise<SignatureProofList[]> {
    const B = msgHash.length
    if (B === 0) return []
    const n = Math.ceil(Math.log2(keys.length)),
        cat = (a: Uint8Array[], w: number) => { const o = new Uint8Array(w * a.length); a.forEach((x, i) => o.set(x, w * i)); return o },
        pks = await Promise.all(publicKey.map(async (k) => new Uint8Array(await crypto.subtle.exportKey('raw', k)))),
        tapes = Array.from({ length: B }, () => proveTape(params.SecLevel, n)),
        res = await nativeFor(params).proveBatch(paramsHandle(params), cat(msgHash, 32), cat(sigBytes, 64), cat(pks, 65), Uint32Array.from(which),
            ringBytes(keys), cat(tapes, tapes[0].length), params.SecLevel)
    return Array.from({ length: B }, (_, b) => readProof(res.proofs.subarray(b * res.stride, b * res.stride + res.lens[b]), params.SecLevel, params.ProofGroup.c))
}
