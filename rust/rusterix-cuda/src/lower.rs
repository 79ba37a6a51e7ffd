//! NodeOp tree -> flat word stream of `rxc_program` (include/rxcuda.h): one word per op, `opcode | a << 8`,
//! `If` / `For` lowered to Jz / Jmp / Mark / Trunc / Unmark, every function body ends with End.
//! Twin of rusterix_b200/vm.py::Program.flatten (which the device VM and the CPU oracle are tested against).
use rusteria::{NodeOp, Program};
use std::collections::HashMap;

pub const MAX_LOCALS: usize = 32;
pub const MAX_GLOBALS: usize = 16;
pub const MAX_CALL_DEPTH: usize = 8;
pub const MAX_LOOP_DEPTH: usize = 8;

// flat-code-only opcodes, after the 90 NodeOp variants (RXVM_JZ .. RXVM_END)
pub const OP_JZ: u32 = 90;
pub const OP_JMP: u32 = 91;
pub const OP_MARK: u32 = 92;
pub const OP_TRUNC: u32 = 93;
pub const OP_UNMARK: u32 = 94;
pub const OP_END: u32 = 95;

pub struct FlatProgram {
    pub words: Vec<u32>,
    pub entry: u32,
    pub shade_locals: u32,
    pub n_globals: u32,
    pub sets_opacity: bool,
}

/// Opcode of a NodeOp = its position in the enum declaration (rusteria/src/node/nodeop.rs:12-103).
pub fn opcode(op: &NodeOp) -> u32 {
    use NodeOp::*;
    match op {
        LoadGlobal(_) => 0, StoreGlobal(_) => 1, LoadLocal(_) => 2, StoreLocal(_) => 3, Swap => 4, GetComponents(_) => 5,
        SetComponents(_) => 6, If(_, _) => 7, For(_, _, _, _) => 8, Push(_) => 9, FunctionCall(_, _, _) => 10, Return => 11,
        Dup => 12, Clear => 13, Pack2 => 14, Pack3 => 15, Add => 16, Sub => 17, Mul => 18, Div => 19, Length => 20,
        Length2 => 21, Length3 => 22, Abs => 23, Sin => 24, Sin1 => 25, Sin2 => 26, Cos => 27, Cos1 => 28, Cos2 => 29,
        Tan => 30, Atan => 31, Atan2 => 32, Rotate2D => 33, Dot => 34, Dot2 => 35, Dot3 => 36, Cross => 37, Normalize => 38,
        Floor => 39, Ceil => 40, Round => 41, Fract => 42, Mod => 43, Degrees => 44, Radians => 45, Min => 46, Max => 47,
        Mix => 48, Smoothstep => 49, Step => 50, Clamp => 51, Sqrt => 52, Pow => 53, Log => 54, Print => 55, Eq => 56,
        Ne => 57, Lt => 58, Le => 59, Gt => 60, Ge => 61, And => 62, Or => 63, Not => 64, Neg => 65, UV => 66, SetUV => 67,
        Normal => 68, SetNormal => 69, Hitpoint => 70, Time => 71, Sample => 72, SampleNormal => 73, Color => 74,
        SetColor => 75, Roughness => 76, SetRoughness => 77, Metallic => 78, SetMetallic => 79, Emissive => 80,
        SetEmissive => 81, Opacity => 82, SetOpacity => 83, Bump => 84, SetBump => 85, Alloc => 86, Iterate => 87, Save => 88,
        PaletteIndex => 89,
    }
}

/// GetComponents operand: entries > 2 are skipped (execution.rs:139-147); no entry or more than three -> 7 (pushes 0).
fn swizzle_get(sw: &[u8]) -> u32 {
    let comps: Vec<u32> = sw.iter().filter(|c| **c <= 2).map(|c| *c as u32).collect();
    if comps.is_empty() || comps.len() > 3 { return 7; }
    let mut a = comps.len() as u32;
    for (i, c) in comps.iter().enumerate() { a |= c << (3 + 2 * i); }
    a
}

/// SetComponents operand: n = swizzle length if 1..=3 else 0; an entry > 2 is encoded as 3 (ignored) (execution.rs:163-168).
fn swizzle_set(sw: &[u8]) -> u32 {
    let n = if (1..=3).contains(&sw.len()) { sw.len() } else { 0 };
    let mut a = n as u32;
    for (i, c) in sw.iter().take(n).enumerate() { a |= (if *c <= 2 { *c as u32 } else { 3 }) << (3 + 2 * i); }
    a
}

fn for_each_call(ops: &[NodeOp], f: &mut dyn FnMut(u8, u8, usize)) {
    for op in ops {
        match op {
            NodeOp::FunctionCall(arity, total, index) => f(*arity, *total, *index),
            NodeOp::If(t, e) => { for_each_call(t, f); if let Some(e) = e { for_each_call(e, f); } }
            NodeOp::For(a, b, c, d) => { for part in [a, b, c, d] { for_each_call(part, f); } }
            _ => {}
        }
    }
}

fn call_depth(p: &Program, f: usize, stack: &mut Vec<usize>) -> Result<usize, String> {
    if stack.contains(&f) { return Err("recursive shader functions are not supported on the device".into()); }
    let body = p.user_functions.get(f).ok_or_else(|| format!("FunctionCall to missing function {f} (the reference panics)"))?;
    stack.push(f);
    let mut depth = 1usize;
    let mut err = None;
    for_each_call(body, &mut |_, _, index| {
        if err.is_none() { match call_depth(p, index, stack) { Ok(d) => depth = depth.max(1 + d), Err(e) => err = Some(e) } }
    });
    stack.pop();
    match err { Some(e) => Err(e), None => Ok(depth) }
}

struct Lowering<'a> { p: &'a Program, words: Vec<u32>, fixups: Vec<(usize, usize)> }

impl<'a> Lowering<'a> {
    fn emit(&mut self, op: u32, a: usize) -> Result<(), String> {
        if a >= (1 << 24) { return Err("operand does not fit 24 bits".into()); }
        self.words.push(op | ((a as u32) << 8));
        Ok(())
    }
    fn patch(&mut self, at: usize) { let target = self.words.len() as u32; self.words[at] |= target << 8; }

    fn lower(&mut self, ops: &[NodeOp], n_locals: usize, loop_depth: usize) -> Result<(), String> {
        for op in ops {
            let code = opcode(op);
            match op {
                NodeOp::Alloc | NodeOp::Iterate | NodeOp::Save => return Err("texture-baking op in shade(): not available on the device".into()),
                NodeOp::LoadLocal(i) | NodeOp::StoreLocal(i) => {
                    if *i >= n_locals { return Err(format!("local {i} outside the frame's {n_locals} locals (the reference panics)")); }
                    self.emit(code, *i)?;
                }
                NodeOp::LoadGlobal(i) | NodeOp::StoreGlobal(i) => {
                    if *i >= self.p.globals { return Err(format!("global {i} outside the program's {} globals (the reference panics)", self.p.globals)); }
                    self.emit(code, *i)?;
                }
                NodeOp::GetComponents(sw) => self.emit(code, swizzle_get(sw) as usize)?,
                NodeOp::SetComponents(sw) => self.emit(code, swizzle_set(sw) as usize)?,
                NodeOp::Push(v) => { self.emit(code, 0)?; self.words.extend([v.x.to_bits(), v.y.to_bits(), v.z.to_bits()]); }
                NodeOp::FunctionCall(arity, total, index) => {
                    if *total as usize > MAX_LOCALS { return Err("too many locals in a function call for the device VM".into()); }
                    self.emit(code, *arity as usize | ((*total as usize) << 8))?;
                    self.fixups.push((self.words.len(), *index));
                    self.words.push(0);
                }
                NodeOp::If(then_ops, else_ops) => {            // execution.rs:286-293
                    self.emit(OP_JZ, 0)?;
                    let jz = self.words.len() - 1;
                    self.lower(then_ops, n_locals, loop_depth)?;
                    if let Some(else_ops) = else_ops {
                        self.emit(OP_JMP, 0)?;
                        let jmp = self.words.len() - 1;
                        self.patch(jz);
                        self.lower(else_ops, n_locals, loop_depth)?;
                        self.patch(jmp);
                    } else {
                        self.patch(jz);
                    }
                }
                NodeOp::For(init, cond, incr, body) => {       // execution.rs:259-285
                    if loop_depth + 1 > MAX_LOOP_DEPTH { return Err("loop nesting exceeds the device limit".into()); }
                    self.emit(OP_MARK, 0)?;
                    self.lower(init, n_locals, loop_depth + 1)?;
                    self.emit(OP_TRUNC, 0)?;
                    let top = self.words.len();
                    self.lower(cond, n_locals, loop_depth + 1)?;
                    self.emit(OP_JZ, 0)?;
                    let jz = self.words.len() - 1;
                    self.emit(OP_TRUNC, 0)?;
                    self.lower(body, n_locals, loop_depth + 1)?;
                    self.emit(OP_TRUNC, 0)?;
                    self.lower(incr, n_locals, loop_depth + 1)?;
                    self.emit(OP_TRUNC, 0)?;
                    self.emit(OP_JMP, top)?;
                    self.patch(jz);
                    self.emit(OP_UNMARK, 0)?;
                }
                _ => self.emit(code, 0)?,
            }
        }
        Ok(())
    }
}

/// `Err` = the program cannot run on the device (the caller falls back to the CPU `rasterize`).
pub fn lower_program(p: &Program) -> Result<FlatProgram, String> {
    let Some(shade) = p.shade_index else {
        return Ok(FlatProgram { words: vec![], entry: 0, shade_locals: 0, n_globals: p.globals as u32, sets_opacity: false });
    };
    if call_depth(p, shade, &mut vec![])? > MAX_CALL_DEPTH { return Err("call depth exceeds the device limit".into()); }
    if p.globals > MAX_GLOBALS || p.shade_locals > MAX_LOCALS { return Err("too many globals/locals for the device VM".into()); }
    // functions reachable from shade(), and the largest frame each is called with
    let (mut order, mut todo, mut n_locals_of) = (Vec::new(), vec![shade], HashMap::new());
    n_locals_of.insert(shade, p.shade_locals);
    while let Some(f) = todo.pop() {
        if order.contains(&f) { continue; }
        order.push(f);
        for_each_call(&p.user_functions[f], &mut |_, total, index| {
            let e = n_locals_of.entry(index).or_insert(0usize);
            *e = (*e).max(total as usize);
            todo.push(index);
        });
    }
    let mut l = Lowering { p, words: Vec::new(), fixups: Vec::new() };
    let mut offsets = HashMap::new();
    for f in &order {
        offsets.insert(*f, l.words.len() as u32);
        l.lower(&p.user_functions[*f], *n_locals_of.get(f).unwrap_or(&0), 0)?;
        l.emit(OP_END, 0)?;
    }
    for (pos, index) in l.fixups.clone() { l.words[pos] = offsets[&index]; }
    Ok(FlatProgram { entry: offsets[&shade], words: l.words, shade_locals: p.shade_locals as u32, n_globals: p.globals as u32,
                     sets_opacity: p.shader_supports_opacity() })
}
