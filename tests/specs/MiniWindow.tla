---------------------------- MODULE MiniWindow ----------------------------
(* Third synthetic Sequences spec: SubSeq with a run-time lower bound, Len of a concatenation, \E over 1 .. Len(s)
   with a run-time index, a record with a tuple-valued field, membership in a product with an infinite component. *)
EXTENDS Integers, Sequences

CONSTANTS N, K

VARIABLES buf, win, note

Items == 1 .. K

TypeOk ==
    /\ buf \in Seq(Items)
    /\ win \in [lo : 1 .. N + 1, pair : Items \X (0 .. N)]
    /\ note \in BOOLEAN

Init ==
    /\ buf = << >>
    /\ win = [lo |-> 1, pair |-> <<1, 0>>]
    /\ note = FALSE

Push(x) ==
    /\ Len(buf) < N
    /\ buf' = buf \o <<x>>
    /\ UNCHANGED <<win, note>>

Slide ==
    /\ win.lo <= Len(buf)
    /\ win' = [lo |-> win.lo + 1, pair |-> <<buf[win.lo], Len(SubSeq(buf, win.lo, Len(buf)))>>]
    /\ note' = (\E i \in 1 .. Len(buf) : i >= win.lo /\ buf[i] = K)
    /\ UNCHANGED buf

Trim ==
    /\ win.lo > 1
    /\ buf' = SubSeq(buf, win.lo, Len(buf))
    /\ win' = [win EXCEPT !.lo = 1]
    /\ note' = (Len(buf \o buf) > N)

Next == (\E x \in Items : Push(x)) \/ Slide \/ Trim

PairOk == win.pair \in Nat \X Nat /\ win.pair[2] <= N
LoOk == win.lo <= Len(buf) + 1
=============================================================================
