---------------------------- MODULE MiniQueue ----------------------------
(* Synthetic spec for the front-end coverage tests (SURVEY section 8f, row 4): module Sequences (Seq, Len, Head,
   Tail, Append, SubSeq, \o, DOMAIN / indexing / EXCEPT on a sequence), tuples as values, a Cartesian product in
   the type invariant, and a RECURSIVE operator.  A bounded FIFO between a producer and a consumer. *)
EXTENDS Integers, Sequences

CONSTANTS Cap, MaxVal

VARIABLES queue,    \* the channel: a sequence of at most Cap values
          last,     \* <<value, seen>>: the value delivered last and whether anything was delivered yet
          total     \* running sum of the values in the channel

Vals == 0 .. MaxVal

RECURSIVE SumSeq(_)
SumSeq(s) == IF s = << >> THEN 0 ELSE Head(s) + SumSeq(Tail(s))

TypeOk ==
    /\ queue \in Seq(Vals)
    /\ Len(queue) <= Cap
    /\ last \in Vals \X BOOLEAN
    /\ total \in 0 .. Cap * MaxVal

Init ==
    /\ queue = << >>
    /\ last = <<0, FALSE>>
    /\ total = 0

Put(v) ==
    /\ Len(queue) < Cap
    /\ queue' = Append(queue, v)
    /\ total' = total + v
    /\ UNCHANGED last

Get ==
    /\ queue # << >>
    /\ last' = <<Head(queue), TRUE>>
    /\ queue' = Tail(queue)
    /\ total' = total - Head(queue)

Bump ==
    \E i \in DOMAIN queue :
        /\ queue[i] < MaxVal
        /\ queue' = [queue EXCEPT ![i] = @ + 1]
        /\ total' = total + 1
        /\ UNCHANGED last

Rotate ==
    /\ Len(queue) >= 2
    /\ queue' = SubSeq(queue, 2, Len(queue)) \o <<queue[1]>>
    /\ UNCHANGED <<last, total>>

Next ==
    \/ \E v \in Vals : Put(v)
    \/ Get
    \/ Bump
    \/ Rotate

SumOk == total = SumSeq(queue)
LastOk == last[2] \/ last[1] = 0
Bounded == Len(queue) <= Cap /\ \A i \in 1 .. Len(queue) : queue[i] \in Vals
=============================================================================
