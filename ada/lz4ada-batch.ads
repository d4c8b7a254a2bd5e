--  LZ4Ada.Batch -- the batched device entry point (new; no counterpart in the reference).
--  UNCOMPILED in this image (no GNAT); the tested implementation of the same algorithm is
--  bo_lz4_ada_b200/csrc/host/batch.cpp (lz4ada_batch_* in include/lz4b200.h).
with Ada.Exceptions;
with Ada.Strings.Unbounded;

package LZ4Ada.Batch is

   --  One input stream: what a caller would feed to Init (For_All) + Update until end of input,
   --  i.e. one or more concatenated modern / legacy / skippable frames.
   type Stream_Span is record
      Src_Off, Src_Len : U64;   --  where the stream lies in Source
      Dst_Off, Dst_Cap : U64;   --  where its output goes in Destination (Dst_Cap = 0: let the planner place it)
   end record;
   type Stream_Spans is array (Positive range <>) of Stream_Span;

   type Outcome is record
      Raised       : Ada.Exceptions.Exception_Id := Ada.Exceptions.Null_Id;
      --  Null_Id, or Checksum_Error / Data_Corruption / Not_Supported / Too_Little_Memory'Identity:
      --  the first exception the serial Update loop would have raised on this stream
      Message      : Ada.Strings.Unbounded.Unbounded_String;   --  exactly the serial message
      End_Of_Frame : LZ4Ada.End_Of_Frame := No;                 --  Is_End_Of_Frame after the last byte
      Dst_Off      : U64 := 0;
      Out_Len      : U64 := 0;                                  --  bytes produced before any error
   end record;
   type Outcomes is array (Positive range <>) of Outcome;

   Device_Error : exception;   --  CUDA failure: never confused with an LZ4 data error

   --  Bytes of Destination the batch needs (upper bound: every block at its frame's block maximum).
   function Output_Bytes (Source : in Octets; Spans : in Stream_Spans;
                          Reservation : in Memory_Reservation := For_All) return U64;

   procedure Decompress (Source      : in     Octets;
                         Spans       : in out Stream_Spans;
                         Destination :    out Octets;
                         Results     :    out Outcomes;
                         Reservation : in     Memory_Reservation := For_All;
                         GPU         : in     Natural := 0)
     with Pre => Results'Length = Spans'Length;

end LZ4Ada.Batch;
