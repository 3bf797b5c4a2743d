#!/usr/bin/env node
// gen_ts_vectors.mjs — pin the oracle to the REAL TypeScript implementation (run where node >= 24 exists).
//
//   git clone https://github.com/cloudflare/zkp-ecdsa && cd zkp-ecdsa && git checkout v0.2.6 && npm ci && npm run build
//   node <repo>/oracle/gen_ts_vectors.mjs /path/to/zkp-ecdsa <repo>/tests/golden/ts_inputs.json <repo>/tests/golden
//
// For every golden case (tests/golden/export_ts_inputs.py) it
//   * replaces crypto.getRandomValues (the only randomness source, src/bignum/big.ts:175) by a reader of the
//     committed tape — the tapes are pre-filtered, so rnd()'s rejection loop never redraws and the reference
//     consumes exactly the draws the C ABI documents (include/zkattest.h);
//   * checks generateParamsList() against h_nist / h_proof;
//   * calls the reference's own proveSignatureList (src/zkpAttestList.ts:104) and verifySignatureList (:147);
//   * writes the proof as the flat concatenation of the reference's toBytes() primitives in class-field order
//     (the layout of include/zkattest.h) to <outdir>/ts_<tag>.bin = B x [u32 little-endian length, bytes], and the
//     verdicts to <outdir>/ts_<tag>.verdict.json.
// tests/test_ts_vectors.py compares these files with the oracle-made fixtures when they are present.
import { readFileSync, writeFileSync } from 'node:fs'
import { webcrypto } from 'node:crypto'
import { pathToFileURL } from 'node:url'
import path from 'node:path'

const [, , refDir, inputsPath, outDir] = process.argv
if (!refDir || !inputsPath || !outDir) {
    console.error('usage: gen_ts_vectors.mjs <zkp-ecdsa checkout (built)> <ts_inputs.json> <outdir>')
    process.exit(2)
}
const hex = (s) => Uint8Array.from(Buffer.from(s, 'hex'))

// ---- tape-backed crypto: subtle stays real (SHA-256, importKey/exportKey), getRandomValues reads the tape
let tape = new Uint8Array(0), pos = 0
function useTape(bytes) { tape = bytes; pos = 0 }
const mock = {
    subtle: webcrypto.subtle,
    getRandomValues(buf) {
        if (pos + buf.length > tape.length) throw new Error(`tape exhausted at ${pos}+${buf.length}`)
        buf.set(tape.subarray(pos, pos + buf.length))
        pos += buf.length
        return buf
    },
}
Object.defineProperty(globalThis, 'crypto', { value: mock, configurable: true, writable: true })

const ref = await import(pathToFileURL(path.join(refDir, 'lib', 'src', 'index.js')).href)
const { generateParamsList, proveSignatureList, verifySignatureList } = ref

// ---- flat writer (include/zkattest.h; mirrors bindings/node/zkpAttestListGpu.ts::writeProof)
function flat(proof) {
    const parts = []
    const np = (p) => { const b = p.toBytes(); parts.push(b.length === 1 ? new Uint8Array(65) : b) }
    const wp = (p) => parts.push(p.toBytes())
    const sc = (s) => parts.push(s.toBytes())
    const mult = (m) => { [m.C_4, m.A_x, m.A_y, m.A_z, m.A_4_1, m.A_4_2].forEach(wp); [m.t_x, m.t_y, m.t_z, m.t_rx, m.t_ry, m.t_rz, m.t_r4].forEach(sc) }
    const eq = (e) => { wp(e.A_1); wp(e.A_2); sc(e.t_x); sc(e.t_r1); sc(e.t_r2) }
    np(proof.R); np(proof.comS1); wp(proof.keyXcom); wp(proof.keyYcom)
    for (const e of proof.expProof) {
        if (e.alpha) {
            parts.push(Uint8Array.of(1)); np(e.A); wp(e.Tx); wp(e.Ty); sc(e.alpha); sc(e.beta1); sc(e.beta2); sc(e.beta3)
        } else {
            parts.push(Uint8Array.of(0)); np(e.A); wp(e.Tx); wp(e.Ty); sc(e.z); sc(e.z2)
            const pa = e.proof
            ;[pa.C_8, pa.C_10, pa.C_11, pa.C_13].forEach(wp)
            ;[pa.pi_8, pa.pi_10, pa.pi_11, pa.pi_13].forEach(mult)
            eq(pa.pi_x); eq(pa.pi_y); sc(e.r1); sc(e.r2)
        }
    }
    const gk = proof.membershipProof
    parts.push(Uint8Array.of(gk.cl.length))
    for (const arr of [gk.cl, gk.ca, gk.cb, gk.cd]) arr.forEach(wp)
    for (const arr of [gk.f, gk.za, gk.zb]) arr.forEach(sc)
    sc(gk.zd)
    return Buffer.concat(parts)
}

const inputs = JSON.parse(readFileSync(inputsPath, 'utf8'))
let failures = 0
for (const [tag, c] of Object.entries(inputs)) {
    useTape(hex(c.params_rnd))
    const params = generateParamsList(c.sec_level)
    const okParams = Buffer.from(params.NistGroup.h.toBytes()).toString('hex') === c.h_nist &&
        Buffer.from(params.ProofGroup.h.toBytes()).toString('hex') === c.h_proof
    const keys = c.ring.map((h) => BigInt('0x' + h))
    const chunks = [], verdicts = []
    for (let b = 0; b < c.B; b++) {
        const pk = await webcrypto.subtle.importKey('raw', hex(c.pk[b]), { name: 'ECDSA', namedCurve: 'P-256' }, true, ['verify'])
        useTape(hex(c.tape[b]))
        const proof = await proveSignatureList(params, hex(c.msg_hash[b]), hex(c.sig[b]), pk, c.which[b], keys)
        const drawsUsed = pos / 32
        const bytes = flat(proof)
        useTape(hex(c.vtape[b]))
        let verdict
        try { verdict = (await verifySignatureList(params, hex(c.msg_hash[b]), keys, proof)) ? 1 : 0 } catch (e) { verdict = `throw: ${e.message}` }
        const len = Buffer.alloc(4)
        len.writeUInt32LE(bytes.length)
        chunks.push(len, bytes)
        verdicts.push(verdict)
        const lenOk = bytes.length === c.proof_len[b], verOk = verdict === c.verdict[b]
        if (!lenOk || !verOk) failures++
        console.log(`${tag}[${b}]: ${bytes.length} bytes (${lenOk ? 'length ok' : 'LENGTH DIFFERS from the oracle'}), ${drawsUsed} draws, verdict ${verdict} (${verOk ? 'ok' : 'DIFFERS'})`)
    }
    if (!okParams) { failures++; console.log(`${tag}: generateParamsList DIFFERS from the fixture`) }
    writeFileSync(path.join(outDir, `ts_${tag}.bin`), Buffer.concat(chunks))
    writeFileSync(path.join(outDir, `ts_${tag}.verdict.json`), JSON.stringify({ params_ok: okParams, verdicts }))
}
console.log(failures ? `${failures} mismatch(es) against the oracle fixtures` : 'all lengths / verdicts agree with the oracle fixtures; now run pytest tests/test_ts_vectors.py for the byte comparison')
process.exit(failures ? 1 : 0)
