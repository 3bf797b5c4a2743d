// zkpAttestListGpu.ts — source-compatible TypeScript host over the N-API shim.
// Drop next to src/zkpAttestList.ts of cloudflare/zkp-ecdsa: same exported names and
// signatures (src/zkpAttestList.ts:88,94,104,147); objects are rebuilt from the flat bytes with
// the reference's own deserializePoint / deserializeScalar so writeJson/readJson/eq keep working.
// NOT compiled in this image (no node / tsc).
import { ExpProof } from './exp/exp.js'
import { GKProof } from './proofGK/gk.js'
import { PedersenParams } from './commit/pedersen.js'
import { PointAddProof } from './exp/pointAdd.js'
import { MultProof } from './commit/mult.js'
import { EqualityProof } from './commit/equality.js'
import { SignatureProofList, SystemParametersList } from './zkpAttestList.js'
import { p256, tomEdwards256 } from './curves/instances.js'
import { rnd, toBytes } from './bignum/big.js'
// eslint-disable-next-line @typescript-eslint/no-require-imports
const native = require('../build/Release/zkattest.node')

const NP = 65, WP = 67, NS = 32, WS = 33

// zka_params_create builds the h tables of a parameter set (7.3 GB of HBM, ~0.2 s with the default
// window widths): one native handle per parameter set, kept for the life of the process.
const handles = new Map<string, unknown>()
function paramsHandle(params: SystemParametersList): unknown {
    const hn = params.NistGroup.h.toBytes(), hp = params.ProofGroup.h.toBytes()
    const key = Buffer.from(hn).toString('hex') + Buffer.from(hp).toString('hex') + ':' + params.SecLevel
    let h = handles.get(key)
    if (h === undefined) { h = native.paramsCreate(hn, hp, params.SecLevel); handles.set(key, h) }
    return h
}

// rnd()'s rejection loop stays on the host (big.ts:171-181): draw k has a fixed modulus.
function proveTape(secLevel: number, n: number): Uint8Array {
    const draws = 3 + 4 * secLevel + 40 * secLevel + 5 * n, out = new Uint8Array(32 * draws)
    for (let k = 0; k < draws; k++) {
        const nist = k === 0 || (k >= 3 && k < 3 + 4 * secLevel && (k - 3) % 4 < 2)
        out.set(toBytes(rnd(nist ? p256.order : tomEdwards256.order), 32), 32 * k)
    }
    return out
}

class Reader {
    o = 0
    constructor(private b: Uint8Array) {}
    take(n: number) { const v = this.b.subarray(this.o, this.o + n); this.o += n; return v }
    np() { return p256.deserializePoint(this.take(NP)) }
    wp() { return tomEdwards256.deserializePoint(this.take(WP)) }
    ns() { return p256.deserializeScalar(this.take(NS)) }
    ws() { return tomEdwards256.deserializeScalar(this.take(WS)) }
}
function readMult(r: Reader) {
    const p = [r.wp(), r.wp(), r.wp(), r.wp(), r.wp(), r.wp()], s = [r.ws(), r.ws(), r.ws(), r.ws(), r.ws(), r.ws(), r.ws()]
    return new MultProof(p[0], p[1], p[2], p[3], p[4], p[5], s[0], s[1], s[2], s[3], s[4], s[5], s[6])
}
function readEq(r: Reader) { return new EqualityProof(r.wp(), r.wp(), r.ws(), r.ws(), r.ws()) }
function readProof(bytes: Uint8Array, secLevel: number): SignatureProofList {
    const r = new Reader(bytes), R = r.np(), comS1 = r.np(), kx = r.wp(), ky = r.wp(), exps: ExpProof[] = []
    for (let i = 0; i < secLevel; i++) {
        const tag = r.take(1)[0], A = r.np(), Tx = r.wp(), Ty = r.wp()
        if (tag === 1) exps.push(new ExpProof(A, Tx, Ty, r.ns(), r.ns(), r.ws(), r.ws()))
        else {
            const z = r.ns(), z2 = r.ns(), c = [r.wp(), r.wp(), r.wp(), r.wp()],
                m = [readMult(r), readMult(r), readMult(r), readMult(r)], ex = readEq(r), ey = readEq(r)
            exps.push(new ExpProof(A, Tx, Ty, undefined, undefined, undefined, undefined, z, z2,
                new PointAddProof(c[0], c[1], c[2], c[3], m[0], m[1], m[2], m[3], ex, ey), r.ws(), r.ws()))
        }
    }
    const n = r.take(1)[0], pts = (k: number) => Array.from({ length: k }, () => r.wp()), scs = (k: number) => Array.from({ length: k }, () => r.ws())
    const gk = new GKProof(pts(n), pts(n), pts(n), pts(n), scs(n), scs(n), scs(n), r.ws())
    return new SignatureProofList(R, comS1, kx, ky, exps, gk)
}

export function generateParamsList(secLevel = 80): SystemParametersList {
    const rnd64 = new Uint8Array(64)
    rnd64.set(toBytes(rnd(p256.order), 32), 0)
    rnd64.set(toBytes(rnd(tomEdwards256.order), 32), 32)
    const { hNist, hProof } = native.paramsGenerate(rnd64)
    return new SystemParametersList(
        new PedersenParams(p256, p256.generator(), p256.deserializePoint(hNist)),
        new PedersenParams(tomEdwards256, tomEdwards256.generator(), tomEdwards256.deserializePoint(hProof)),
        secLevel)
}

export async function proveSignatureList(params: SystemParametersList, msgHash: Uint8Array, sigBytes: Uint8Array,
    publicKey: CryptoKey, which: number, keys: bigint[]): Promise<SignatureProofList> {
    const pk = new Uint8Array(await crypto.subtle.exportKey('raw', publicKey)),       // zkpAttestList.ts:113
        ring = new Uint8Array(32 * keys.length)
    keys.forEach((k, i) => ring.set(toBytes(((k % tomEdwards256.order) + tomEdwards256.order) % tomEdwards256.order, 32), 32 * i))
    const n = Math.ceil(Math.log2(keys.length)),
        handle = paramsHandle(params),
        res = await native.proveBatch(handle, msgHash, sigBytes, pk, Uint32Array.of(which), ring,
            proveTape(params.SecLevel, n), params.SecLevel)
    return readProof(res.proofs.subarray(0, res.lens[0]), params.SecLevel)
}
// verifySignatureList(params, msgHash, keys, proof): serialise `proof` with the reference's
// toBytes() in the same field order (writer mirrors Reader above), call native.verifyBatch,
// return ok[0] !== 0; status != 0 is rethrown as Error(message).
